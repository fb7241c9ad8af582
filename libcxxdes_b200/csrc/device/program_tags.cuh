// device/program_tags.cuh -- token tags of the process-program interpreters (program_kernel.cu, program_fast.cu).
//
// A token names what it resumes: a process (coro_data) and/or an any_of/all_of handler (token.ipp:29-32).
// Programs have at most 254 processes, so the 32-bit tag keeps 8 bits for the process (0xFF = null coro_data)
// and 24 bits for the handler, which is named by the SERIAL NUMBER of the composition that created it
// (serial + 1; 0 = null handler), not by a slot: a handler that has fired can leave tokens behind -- the loser
// of any_of(p, timeout(t)) -- and they must keep missing whatever reuses its storage.  A replication may
// therefore run 2^24 - 2 compositions; beyond that it reports REP_BAD_PROGRAM (never a silently wrong match).
#pragma once
#include <stdint.h>

namespace cxb {

constexpr uint32_t PTAG_NONE = 0xFFu;
constexpr uint32_t PTAG_MAX_SERIAL = 0xFFFFFEu;

__device__ __forceinline__ uint32_t ptag(uint32_t target, uint32_t serial_plus1 = 0) {
    return (target & 0xFFu) | (serial_plus1 << 8);
}
__device__ __forceinline__ uint32_t ptag_target(uint32_t tag) { return tag & 0xFFu; }
__device__ __forceinline__ uint32_t ptag_serial_plus1(uint32_t tag) { return tag >> 8; }

}  // namespace cxb
