#!/usr/bin/env python
"""The harness' counterpart of the reference's examples/RpgoReadG2o.cpp, PCM part only (the GNC/LM solve stays on GTSAM):

    python tools/rpgo_read_g2o.py <2d|3d> <file.g2o> <pcm_trans_or_odom_thr> <pcm_rot_or_lc_thr> <output_folder> [--pcm]

Like the reference example it loads the whole g2o graph, hands it to the outlier-rejection stage in one update() with
PcmSimple2D/3D parameters (examples/RpgoReadG2o.cpp:129-148; --pcm selects the covariance-based Pcm2D/3D the README
describes) and logs into <output_folder>: result.g2o (values + the factors PCM kept, the layout of writeG2o,
Logger.cpp:64-163), outlier_rejection_status.txt, rpgo_status.csv and the per-group adjacency matrices.
Needs the CUDA library (no CPU fallback)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv):
    if len(argv) < 6 or argv[1] not in ("2d", "3d"):
        print(__doc__)
        return 2
    pkg = importlib.import_module("kimera-rpgo_b200")
    g2o = importlib.import_module("kimera-rpgo_b200.g2o")
    d = 3 if argv[1] == "3d" else 2
    t1, t2 = float(argv[3]), float(argv[4])
    out = argv[5]
    use_pcm = "--pcm" in argv[6:]
    values, edges = (g2o.load3d if d == 3 else g2o.load2d)(argv[2])
    if use_pcm:
        pcm = pkg.PcmGpu(d, pkg.MODE_PCM, odom_threshold=t1, lc_threshold=t2)
    else:  # setPcmSimple3DParams(trans, rot): the same pair for the odometry and the loop check (SolverParams.h)
        pcm = pkg.PcmGpu(d, pkg.MODE_SIMPLE, odom_trans=t1, odom_rot=t2, dist_trans=t1, dist_rot=t2)
    pcm.log_output(out)
    factors = [(pkg.BETWEEN, k1, k2, p, c) for k1, k2, p, c in edges]
    pcm.update(factors, values)
    # the factors of the rebuilt graph in the reference's order (odometry, then the consistent closures of every group);
    # an id can appear twice: Pcm.h:865-869 consumes FMC's scratch buffer, duplicates included
    kept_edges = [edges[i] for i in pcm.output_ids().tolist()]
    g2o.write_g2o(os.path.join(out, "result.g2o"), values, kept_edges, d=d)
    print("%d vertices, %d edges read; %d odometry, %d loop closures, %d inliers; %d factors written to %s"
          % (len(values), len(edges), pcm.num_odom(), pcm.num_lc(), pcm.num_inliers(), len(kept_edges),
             os.path.join(out, "result.g2o")))
    pcm.close()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
