// comm.h — NCCL communicator + global split table of a context (comm.cpp).
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "engine.h"

namespace qw {

struct NcclUniqueId { char internal[128]; };  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128)

struct Comm {
  void* comm = nullptr;  // ncclComm_t
  int rank = 0, world = 1, device = 0;
  std::vector<std::string> split_ids;  // sorted: position = global tie-break rank of the split id
  // host-staged exchange (comm_allgather_host): device staging + its stream
  void* stream = nullptr;
  uint8_t* d_stage = nullptr;
  size_t stage_cap = 0;
};

void comm_unique_id(uint8_t out[128]);
Comm* comm_create(int device, const uint8_t id[128], int rank, int world);
void comm_destroy(Comm* c);
void comm_set_split_table(Comm* c, uint32_t n, const char* const* split_ids);
int comm_split_rank(const Comm* c, const std::string& split_id);  // -1: not in the table
int comm_allgather(void* comm, const void* send, void* recv, size_t bytes, void* stream);
// host buffers: `bytes` from every rank -> recv[world * bytes] on every rank (H2D, ncclAllGather, D2H; blocking).
// Every rank must call it in the same order relative to the communicator's other collectives.
void comm_allgather_host(Comm* c, const uint8_t* send, uint8_t* recv, size_t bytes);

}  // namespace qw
