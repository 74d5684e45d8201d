// comm.hpp -- the host-side rendezvous of the multi-GPU paths (internal).
//
// The reference has no multi-GPU code; BASELINE.json asks for row shards over the
// GPUs of one NVSwitch box.  What the ranks must exchange on the HOST is tiny and happens at set-up
// only: shard bounds, an ncclUniqueId, and the file descriptors of the cuMem allocations that back
// the symmetric (peer-mapped / multicast-bound) rank vectors.  Two implementations of one interface:
//
//   SocketComm  one PROCESS per GPU (torchrun, mpirun, a C launcher): a hub on rank 0 behind an
//               abstract unix-domain socket named after the session string; descriptors travel as
//               SCM_RIGHTS ancillary data.  One box only -- exactly the NVSwitch domain.
//   ThreadComm  one THREAD per GPU inside one process (spmv_b200_pagerank_multi): shared memory,
//               a generation barrier, descriptors passed by dup().
//
// No CUDA in here: the whole file is covered by CPU tests (tests/test_comm_cpu.py).
#pragma once

#include <cstddef>

namespace spmv {
namespace b200 {

class Comm {
public:
    virtual ~Comm() {}
    int rank() const { return rank_; }
    int world() const { return world_; }
    // recv holds world * bytes; slot p is what rank p sent.  0 on success.
    virtual int allgather(const void* send, void* recv, size_t bytes) = 0;
    virtual int barrier() = 0;
    // every rank contributes one open descriptor; fds_out[p] is a NEW descriptor (owned by the
    // caller) that refers to rank p's open file description.  0 on success.
    virtual int allgather_fds(int my_fd, int* fds_out) = 0;
protected:
    int rank_ = 0, world_ = 1;
};

// one process per rank; `session` names the rendezvous (same string on every rank, unique per job)
int comm_create_socket(int rank, int world, const char* session, int timeout_s, Comm** out);

// one thread per rank: the group hands out world communicators that share state
class ThreadCommGroup;
ThreadCommGroup* thread_comm_group_create(int world);
Comm* thread_comm_get(ThreadCommGroup* group, int rank);  // owned by the group
void thread_comm_group_destroy(ThreadCommGroup* group);

}  // namespace b200
}  // namespace spmv
