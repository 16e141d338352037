// mmm_cells.cu — cutoff mode (opt-in, mmm_set_cutoff(rc > 0)): cell-list build and the pair
// kernel over neighbouring cells.  The reference itself never sets a cutoff (NoCutoff), so this
// path is an extension whose oracle applies the same truncation.
//
// Round-1 status: not built yet.  The entry points exist so the ABI is complete; they report
// MMM_ERR_STATE instead of silently falling back to anything.
#include "mmm_internal.cuh"

int64_t mmm_cells_energy_slots(const mmm_system* h) { (void)h; return 1; }

int mmm_launch_pair_cutoff(mmm_system* h, const int* d_skip) {
  (void)d_skip;
  return mmm_fail(h, MMM_ERR_STATE, "cutoff mode is not available in this build; use mmm_set_cutoff(h, 0)");
}

extern "C" int mmm_get_cell_list(mmm_handle h, int32_t* order_out, uint32_t* key_out) {
  (void)order_out;
  (void)key_out;
  if (!h) return MMM_ERR_ARG;
  return mmm_fail(h, MMM_ERR_STATE, "cutoff mode is not available in this build");
}
