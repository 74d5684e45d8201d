// symm.hpp -- symmetric (peer-mapped, optionally multicast-bound) device buffers (internal).
#pragma once

#include "comm.hpp"

#include <cstddef>

namespace spmv {
namespace b200 {

constexpr int kSymmMaxRanks = 8;  // GPUs of one NVSwitch box

struct SymmBuffer {
    void* local = nullptr;                 // this rank's buffer (== peer[rank])
    void* peer[kSymmMaxRanks] = {};        // rank p's buffer mapped here (NVLink loads / stores)
    void* mc = nullptr;                    // multicast mapping: a multimem.st reaches every rank's buffer
    size_t bytes = 0, mapped_bytes = 0;
    int world = 0, rank = 0, device = 0;
    unsigned long long handles[kSymmMaxRanks] = {};  // CUmemGenericAllocationHandle of every mapping
    unsigned long long mc_handle = 0;
    bool mc_bound = false;
};

bool symm_supported();
// collective over comm; 0 on success on EVERY rank, -1 on every rank otherwise
int symm_alloc(Comm* comm, size_t bytes, bool want_multicast, SymmBuffer* out);
void symm_free(SymmBuffer* b);

}  // namespace b200
}  // namespace spmv
