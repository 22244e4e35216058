// ppr_internal.h -- declarations shared by the host (.cc) and device (.cu) halves of libppr_b200.so
#ifndef PPRB200_INTERNAL_H
#define PPRB200_INTERNAL_H

#include <cstdint>
#include <functional>
#include <vector>

#include "../../include/pprb200.h"

namespace pprb200 {

// records the message for pprb200_last_error() and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

int validate_csr(const int64_t* row_ptr, const int32_t* col, int32_t n);
// device_component (optional): explores the whole component of `root` on the device -- sets seen[v] = 1 and colour[v] =
// parity of the BFS distance from root for its nodes and returns true, or returns false having touched nothing
using ComponentFn = std::function<bool(int32_t root, uint8_t* seen, uint8_t* colour)>;
int find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour, const ComponentFn* device_component = nullptr);

// ---- host-side parallelism (the front half of a call: colouring, storage order, CSR encode) ----
// Worker threads are created once per process and reused (PPRB200_HOST_THREADS, default min(hardware threads, 16)).
// host_parallel(parts, fn) runs fn(part) for part = 0..parts-1 on the pool (the caller takes part in the work) and
// returns when all are done; parts <= host_threads(). Calls are serialised.
int host_threads();
void host_parallel(int parts, const std::function<void(int)>& fn);
// splits [0, n) into `parts` ranges and runs fn(part, lo, hi); runs inline when n is small or parts == 1
void host_parallel_for(int64_t n, int64_t grain, const std::function<void(int, int64_t, int64_t)>& fn);
// CSR transpose (predecessor lists by ascending source id, multiplicity kept), parallel; edges whose source has
// skip_source[u] != 0 are left out
void host_transpose(const int64_t* row_ptr, const int32_t* col, int32_t n, std::vector<int64_t>& prow, std::vector<int32_t>& pcol,
                    const uint8_t* skip_source = nullptr);
// in-degree of every node (multiplicity counted), parallel
void host_indegree(const int32_t* col, int64_t e, int32_t n, uint32_t* indeg);

}  // namespace pprb200

#endif
