// tests/emu/emu_radau.cc — the batched Radau kernel (csrc/radau_batch.cu), compiled for the host and run by the SIMT
// emulator: one CTA of four warps works through all columns of the queue.  TEST INFRASTRUCTURE ONLY (simt_emu.h).
#include <cstdint>
#include <vector>

#include "cuda_runtime.h"
#include "simt_emu.h"

#include "../../integrating-diagenetic-equations-using-python_b200/csrc/radau_batch.cu"

static int run_radau(bool team, double* y, const marlpde_column_params* params, marlpde_column_state* state, int n_columns,
                     int n_cells, const marlpde_rk45_options* opt, const double* t_eval, double* snap, int64_t* stats,
                     int32_t* ev_counts, double* ev_times) {
  using namespace marlpde;
  std::vector<double> work(rd::work_doubles(n_cells) * (size_t)n_columns + 16, 0.0);
  int32_t queue = 0;
  rd::Args a;
  a.g_y = y;
  a.g_params = params;
  a.g_state = state;
  a.g_t_eval = t_eval;
  a.g_snap = snap;
  a.g_stats = stats;
  a.g_ev_counts = ev_counts;
  a.g_ev_times = ev_times;
  a.g_work = work.data();
  a.g_queue = &queue;
  a.n_columns = n_columns;
  a.N = n_cells;
  a.opt = *opt;
  const bool vd = (a.opt.flags & MARLPDE_FLAG_VAR_DPHI) != 0, fd = (a.opt.flags & MARLPDE_FLAG_JAC_FD) != 0;
  if (team)   // one CTA of two warps per column (the latency shape); one CTA works through the queue here
    return simt::run_block(64, 0, [&]() {
      if (vd) rd::radau_kernel<true, false, 2>(a);
      else rd::radau_kernel<false, false, 2>(a);
    });
  return simt::run_block(rd::kWarpsPerCta * 32, 0, [&]() {
    if (vd && fd) rd::radau_kernel<true, true, 1>(a);
    else if (vd) rd::radau_kernel<true, false, 1>(a);
    else if (fd) rd::radau_kernel<false, true, 1>(a);
    else rd::radau_kernel<false, false, 1>(a);
  });
}

extern "C" int emu_radau(double* y, const marlpde_column_params* params, marlpde_column_state* state, int n_columns,
                         int n_cells, const marlpde_rk45_options* opt, const double* t_eval, double* snap,
                         int64_t* stats, int32_t* ev_counts, double* ev_times) {
  return run_radau(false, y, params, state, n_columns, n_cells, opt, t_eval, snap, stats, ev_counts, ev_times);
}

extern "C" int emu_radau_team(double* y, const marlpde_column_params* params, marlpde_column_state* state, int n_columns,
                              int n_cells, const marlpde_rk45_options* opt, const double* t_eval, double* snap,
                              int64_t* stats, int32_t* ev_counts, double* ev_times) {
  return run_radau(true, y, params, state, n_columns, n_cells, opt, t_eval, snap, stats, ev_counts, ev_times);
}
