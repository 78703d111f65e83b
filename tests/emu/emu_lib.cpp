// TEST INFRASTRUCTURE ONLY (see host_backend.hpp): the same C ABI over the host stand-in.
#include "host_backend.hpp"
#define PD_BACKEND pd::HostBackend
#include "../../pulser_diff_b200/csrc/cabi_impl.hpp"

extern "C" int pd_emu_set_segment_budget(pd_plan* p, uint64_t bytes) {
  if (!p) return PD_ERR_INVALID;
  p->eng.bk.segment_budget = (size_t)bytes;
  return PD_OK;
}
