/* comm.h — internal interface of the NCCL data plane (comm.cu). */
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>

#include <string>

namespace rpgo {

struct Comm;

int comm_unique_id(void* id_out_128, std::string* err);
int comm_create(Comm** out, const void* id128, int rank, int world, std::string* err);
void comm_destroy(Comm* c);
const char* comm_error(const Comm* c);
int comm_rank(const Comm* c);
int comm_world(const Comm* c);
int comm_nccl_version();

/* all-gather of the adjacency row chunks {r, 2*world-1-r} of every rank r, in place, on stream st */
int comm_allgather_row_chunks(Comm* c, void* bits, size_t chunk_bytes, cudaStream_t st);
/* device-resident all-reduce / broadcast on stream st (asynchronous) */
int comm_allreduce_i64_device(Comm* c, long long* dev, size_t count, bool is_max, cudaStream_t st);
int comm_bcast_device(Comm* c, void* dev, size_t bytes, int root, cudaStream_t st);
/* host-buffer forms (blocking; pinned staging inside) */
int comm_allreduce_i64_host(Comm* c, long long* host, size_t count, bool is_max, cudaStream_t st);
int comm_bcast_host(Comm* c, void* host, size_t bytes, int root, cudaStream_t st);

}  // namespace rpgo
