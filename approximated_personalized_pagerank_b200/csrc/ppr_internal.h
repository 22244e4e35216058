// ppr_internal.h -- declarations shared by the host (.cc) and device (.cu) halves of libppr_b200.so
#ifndef PPRB200_INTERNAL_H
#define PPRB200_INTERNAL_H

#include <cstdint>

#include "../../include/pprb200.h"

namespace pprb200 {

// records the message for pprb200_last_error() and returns `code`
int fail(int code, const char* fmt, ...) __attribute__((format(printf, 2, 3)));

int validate_csr(const int64_t* row_ptr, const int32_t* col, int32_t n);
int find_partitions(const int64_t* row_ptr, const int32_t* col, int32_t n, uint8_t* colour);

}  // namespace pprb200

#endif
