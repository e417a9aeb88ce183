// tests/emu/emu_rk45.cc — the on-chip RK45 kernel (csrc/rk45_persistent.cu), compiled for the host
// and run one thread block at a time by the SIMT emulator.  TEST INFRASTRUCTURE ONLY: checks kernel control logic
// (indexing, barriers, halo exchange, slot service, event handling) without a GPU; see simt_emu.h.
#include <cstdint>
#include <vector>

#include "cuda_runtime.h"
#include "simt_emu.h"

#include "../../integrating-diagenetic-equations-using-python_b200/csrc/rk45_persistent.cu"

extern "C" int emu_rk45(int variant, double* y, const marlpde_column_params* params, marlpde_column_state* state,
                        int n_columns, int n_cells, const marlpde_rk45_options* opt, const double* t_eval, double* snap,
                        int32_t* ev_counts, double* ev_times) {
  using namespace marlpde;
  std::vector<int32_t> queue_words(1 + 2 * (size_t)(n_columns > 0 ? n_columns : 0), 0);   // counter + lock words + attempt counters
  int32_t& queue = queue_words[0];
  const int budget = 227 * 1024;
  // EMU_GRID blocks, one after the other, on the same column queue (the claim policy sees gridDim.x = EMU_GRID)
  const char* ge = std::getenv("EMU_GRID");
  const int grid = ge && std::atoi(ge) > 0 ? std::atoi(ge) : 1;
  (void)variant;
  Rk45Args a;
  a.g_y = y;
  a.g_params = params;
  a.g_state = state;
  a.g_t_eval = t_eval;
  a.g_snap = snap;
  a.g_queue = &queue;
  a.g_ev_counts = ev_counts;
  a.g_ev_times = ev_times;
  a.n_columns = n_columns;
  a.N = n_cells;
  a.C = columns_per_cta_t<kRk45Threads>(n_cells, budget);
  if (a.C <= 0) return -2;
  const int Hc = (n_cells + 1) / 2;
  a.logG = group_log2(Hc);
  a.warp_perm = ~0ull;
  a.opt = *opt;
  a.n_quanta = 1;
  a.n_whole = 0;
  a.quantum = 0;
  // EMU_SLOTS: pretend the launch has this many resident slots when the library chooses the quanta itself
  const char* se = std::getenv("EMU_SLOTS");
  choose_quanta(a, se && std::atoi(se) > 0 ? std::atoi(se) : grid * a.C);
  const int threads = ((a.C * Hc + 31) / 32) * 32;
  for (int b = 0; b < grid; ++b)
    if (int rc = simt::run_block(threads, Smem<kRk45Threads>::total(a.C),
                                 [&]() {
                                   if (opt->flags & MARLPDE_FLAG_VAR_DPHI) rk45_persistent_kernel<kRk45Threads, true>(a);
                                   else rk45_persistent_kernel<kRk45Threads, false>(a);
                                 }, b, grid))
      return rc;
  return 0;
}
