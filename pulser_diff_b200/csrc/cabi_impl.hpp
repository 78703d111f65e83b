// C-ABI entry points (include/pulser_diff_b200.h), templated on the backend through PD_BACKEND.
// Included by cabi.cu (CUDA, the product) and by tests/emu/emu_lib.cpp (host stand-in, tests only).
#pragma once
#include "engine.hpp"

#ifndef PD_BACKEND
#error "define PD_BACKEND before including cabi_impl.hpp"
#endif

using PdEngine = pd::Engine<PD_BACKEND>;

struct pd_plan {
  PdEngine eng;
  pd_plan(int nq, int batch, int kind, int dev) : eng(nq, batch, kind, dev) {}
};
struct pd_tape {
  pd::Tape tape;
  std::vector<pd::Tape> units;   // pd_evolve_forward_units: one per parameter set
};

namespace {
thread_local std::string g_last_error;
template <class F>
int guarded(F&& f) {
  try {
    f();
    return PD_OK;
  } catch (const pd::Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return PD_ERR_STATE;
  }
}
// Every entry point that launches or copies runs with the plan's device current and restores the caller's
// device afterwards (a plan on cuda:k must not change the device torch sees, nor launch on a foreign one).
struct DeviceScope {
  int prev;
  explicit DeviceScope(int dev) : prev(PD_BACKEND::push_device(dev)) {}
  ~DeviceScope() { PD_BACKEND::pop_device(prev); }
};
template <class F>
int guarded_on(const pd_plan* p, F&& f) {
  if (!p) return guarded(f);
  DeviceScope scope(p->eng.bk.device);
  return guarded(f);
}
void need(bool ok, const char* what) {
  if (!ok) throw pd::Error(PD_ERR_INVALID, what);
}
}  // namespace

extern "C" {

int pd_abi_version(void) { return PD_ABI_VERSION; }
const char* pd_last_error(void) { return g_last_error.c_str(); }
int pd_is_cuda(void) { return PD_BACKEND::is_cuda ? 1 : 0; }
int pd_amplitude_bytes(void) { return (int)sizeof(pd::amp_t); }

void pd_options_default(pd_options* o) {
  if (!o) return;
  std::memset(o, 0, sizeof(*o));
  o->atol = 1e-8;
  o->rtol = 1e-6;
  o->max_steps = 100000;
  o->safety_factor = 0.9;
  o->min_factor = 0.2;
  o->max_factor = 5.0;
  o->max_krylov = 80;
  o->exp_tolerance = 1e-10;
  o->norm_tolerance = 1e-10;
}

int pd_plan_create(pd_plan** out, int32_t n_qubits, int32_t batch, int32_t kind, int32_t device) {
  return guarded([&] {
    need(out != nullptr, "pd_plan_create: out is NULL");
    DeviceScope scope(device);
    *out = new pd_plan(n_qubits, batch, kind, device);
  });
}
int pd_plan_destroy(pd_plan* p) {
  return guarded_on(p, [&] { delete p; });
}
int pd_plan_set_interaction(pd_plan* p, const double* pair_u_host, void* stream) {
  return guarded_on(p, [&] {
    need(p && pair_u_host, "pd_plan_set_interaction: NULL argument");
    p->eng.set_interaction(pair_u_host, stream);
  });
}
int pd_plan_set_terms(pd_plan* p, int32_t n_samples, double dt, int32_t n_det,
                      const uint64_t* det_masks, const double* det_values, int32_t n_amp,
                      const uint64_t* amp_masks, const double* amp_values) {
  return guarded_on(p, [&] {
    need(p != nullptr, "pd_plan_set_terms: plan is NULL");
    need(n_det >= 0 && n_amp >= 0, "pd_plan_set_terms: negative term count");
    need(n_det == 0 || (det_masks && det_values), "pd_plan_set_terms: det arrays missing");
    need(n_amp == 0 || (amp_masks && amp_values), "pd_plan_set_terms: amp arrays missing");
    p->eng.set_terms(n_samples, dt, n_det, det_masks, det_values, n_amp, amp_masks, amp_values);
  });
}
int pd_plan_set_collapse(pd_plan* p, int32_t n_ops, const double* ops_host) {
  return guarded_on(p, [&] {
    need(p != nullptr && n_ops >= 0 && (n_ops == 0 || ops_host), "pd_plan_set_collapse: bad argument");
    p->eng.set_collapse(n_ops, ops_host);
  });
}
int pd_plan_set_path(pd_plan* p, int32_t path) {
  return guarded_on(p, [&] {
    need(p != nullptr && path >= 0 && path <= 5, "pd_plan_set_path: bad argument");
    p->eng.bk.path = path;
  });
}
int pd_hpsi(pd_plan* p, void* stream, double t, const void* in_dev, void* out_dev) {
  return guarded_on(p, [&] {
    need(p && in_dev && out_dev, "pd_hpsi: NULL argument");
    need(in_dev != out_dev, "pd_hpsi: in-place application is not supported");
    p->eng.apply((pd::amp_t*)out_dev, (const pd::amp_t*)in_dev, t, 2, stream);
  });
}
int pd_rhs(pd_plan* p, void* stream, double t, const void* in_dev, void* out_dev) {
  return guarded_on(p, [&] {
    need(p && in_dev && out_dev, "pd_rhs: NULL argument");
    need(in_dev != out_dev, "pd_rhs: in-place application is not supported");
    p->eng.apply((pd::amp_t*)out_dev, (const pd::amp_t*)in_dev, t, 0, stream);
  });
}
int pd_evolve_forward(pd_plan* p, void* stream, int32_t solver, const pd_options* opt,
                      const void* state0_dev, const double* tsave_host, int32_t n_t,
                      void* states_dev, pd_tape** tape_out) {
  return guarded_on(p, [&] {
    need(p && state0_dev && tsave_host && states_dev, "pd_evolve_forward: NULL argument");
    pd_options o;
    if (opt) o = *opt; else pd_options_default(&o);
    need(o.atol > 0 && o.rtol >= 0 && o.max_steps > 0, "pd_evolve_forward: bad tolerances");
    p->eng.bk.path = o.path;
    pd_tape* tp = tape_out ? new pd_tape() : nullptr;
    try {
      p->eng.forward(solver, o, (const pd::amp_t*)state0_dev, tsave_host, n_t,
                     (pd::amp_t*)states_dev, tp ? &tp->tape : nullptr, stream);
    } catch (...) {
      delete tp;
      throw;
    }
    if (tape_out) *tape_out = tp;
  });
}
int pd_evolve_backward(pd_plan* p, void* stream, pd_tape* tape, const void* states_dev,
                       const void* grad_states_dev, double* grad_det_host, double* grad_amp_host,
                       double* grad_pair_u_host, double* grad_tsave_host, void* grad_state0_dev) {
  return guarded_on(p, [&] {
    need(p && tape && states_dev, "pd_evolve_backward: NULL argument");
    p->eng.backward(tape->tape, (const pd::amp_t*)states_dev, (const pd::amp_t*)grad_states_dev,
                    grad_det_host, grad_amp_host, grad_pair_u_host, grad_tsave_host,
                    (pd::amp_t*)grad_state0_dev, stream);
  });
}
int pd_evolve_forward_units(pd_plan* p, void* stream, const pd_options* opt, int32_t n_units,
                            const void* state0_dev, const double* tsave_host, int32_t n_t,
                            const double* det_values_host, const double* amp_values_host,
                            void* states_dev, pd_tape** tape_out) {
  return guarded_on(p, [&] {
    need(p && state0_dev && tsave_host && states_dev && n_units >= 1, "pd_evolve_forward_units: bad argument");
    need((det_values_host || p->eng.prog.n_det() == 0) && (amp_values_host || p->eng.prog.n_amp() == 0),
         "pd_evolve_forward_units: missing coefficient tables");
    pd_options o;
    if (opt) o = *opt; else pd_options_default(&o);
    p->eng.bk.path = o.path;
    pd_tape* tp = tape_out ? new pd_tape() : nullptr;
    try {
      p->eng.forward_units(o, n_units, (const pd::amp_t*)state0_dev, tsave_host, n_t, det_values_host,
                           amp_values_host, (pd::amp_t*)states_dev, tp ? &tp->units : nullptr, nullptr, stream);
    } catch (...) {
      delete tp;
      throw;
    }
    if (tape_out) *tape_out = tp;
  });
}
int pd_evolve_backward_units(pd_plan* p, void* stream, pd_tape* tape, const void* states_dev,
                             const void* grad_states_dev, const double* det_values_host,
                             const double* amp_values_host, double* grad_det_host, double* grad_amp_host,
                             void* grad_state0_dev) {
  return guarded_on(p, [&] {
    need(p && tape && states_dev && !tape->units.empty(), "pd_evolve_backward_units: bad argument");
    p->eng.states_for_fallback_ = (const pd::amp_t*)states_dev;
    try {
      p->eng.backward_units(tape->units, det_values_host, amp_values_host, (const pd::amp_t*)grad_states_dev,
                            grad_det_host, grad_amp_host, (pd::amp_t*)grad_state0_dev, stream);
    } catch (...) {
      p->eng.states_for_fallback_ = nullptr;
      throw;
    }
    p->eng.states_for_fallback_ = nullptr;
  });
}
int64_t pd_tape_unit_steps(const pd_tape* t, int32_t unit, int32_t* attempts_out) {
  if (!t || unit < 0 || (size_t)unit >= t->units.size()) return -1;
  const pd::Tape& u = t->units[unit];
  if (u.n_acc_dev >= 0) {
    if (attempts_out) *attempts_out = u.n_att_dev;
    return u.n_acc_dev;
  }
  if (attempts_out) *attempts_out = (int32_t)u.records.size();
  return (int64_t)u.steps.size();
}
int64_t pd_tape_n_records(const pd_tape* t) { return t ? (int64_t)t->tape.records.size() : 0; }
int pd_tape_records(const pd_tape* t, pd_step_record* out, int64_t capacity) {
  return guarded([&] {
    need(t && out, "pd_tape_records: NULL argument");
    int64_t n = std::min<int64_t>(capacity, (int64_t)t->tape.records.size());
    std::copy(t->tape.records.begin(), t->tape.records.begin() + n, out);
  });
}
int pd_tape_destroy(pd_tape* t) {
  return guarded([&] { delete t; });
}
int pd_expect_diag(pd_plan* p, void* stream, const void* states_dev, int32_t n_t,
                   const double* obs_dev, double* out_host) {
  return guarded_on(p, [&] {
    need(p && states_dev && obs_dev && out_host && n_t >= 1, "pd_expect_diag: bad argument");
    p->eng.expect_diag((const pd::amp_t*)states_dev, n_t, obs_dev, out_host, stream);
  });
}
int pd_rhs_vjp(pd_plan* p, void* stream, double t, const void* state_dev, const void* cot_dev,
               void* grad_state_dev, double* grad_det_host, double* grad_amp_host,
               double* grad_pair_host, double* grad_t_host, int32_t defer_pair) {
  return guarded_on(p, [&] {
    need(p && state_dev && cot_dev, "pd_rhs_vjp: NULL argument");
    double tb = p->eng.rhs_vjp(t, (const pd::amp_t*)state_dev, (const pd::amp_t*)cot_dev,
                               (pd::amp_t*)grad_state_dev, grad_det_host, grad_amp_host,
                               grad_pair_host, defer_pair != 0, stream);
    if (grad_t_host) *grad_t_host = tb;
  });
}
int pd_pair_gradient_flush(pd_plan* p, void* stream, double* grad_pair_host) {
  return guarded_on(p, [&] {
    need(p && grad_pair_host, "pd_pair_gradient_flush: NULL argument");
    p->eng.pair_gradient_flush(grad_pair_host, stream);
  });
}
int pd_lincomb(pd_plan* p, void* stream, void* out_dev, int32_t n_in, const void* const* ins_dev,
               const double* w_host) {
  return guarded_on(p, [&] {
    need(p && out_dev && ins_dev && w_host && n_in >= 1 && n_in <= 8, "pd_lincomb: bad argument");
    for (int j = 0; j < n_in; ++j) need(ins_dev[j] != nullptr, "pd_lincomb: NULL input");
    p->eng.lincomb((pd::amp_t*)out_dev, n_in, (const pd::amp_t* const*)ins_dev, w_host, stream);
  });
}
int pd_dp5_error_sumsq(pd_plan* p, void* stream, const void* const* k_dev, const double* ew_host,
                       const void* y0_dev, const void* y1_dev, double atol, double rtol,
                       double* sumsq_host) {
  return guarded_on(p, [&] {
    need(p && k_dev && ew_host && y0_dev && y1_dev && sumsq_host, "pd_dp5_error_sumsq: NULL argument");
    for (int j = 0; j < 7; ++j)
      need(k_dev[j] != nullptr || ew_host[j] == 0.0, "pd_dp5_error_sumsq: NULL slope with a non-zero weight");
    const pd::amp_t* k[7];
    for (int j = 0; j < 7; ++j) k[j] = k_dev[j] ? (const pd::amp_t*)k_dev[j] : (const pd::amp_t*)y0_dev;
    p->eng.error_sumsq(k, ew_host, (const pd::amp_t*)y0_dev, (const pd::amp_t*)y1_dev, atol, rtol,
                       sumsq_host, stream);
  });
}
int pd_sharded_accumulate(pd_plan* p, void* stream, void* out_dev, const void* psi_dev, double shift,
                          int32_t n_peers, const void* const* peer_slices,
                          const double* coef_host) {
  return guarded_on(p, [&] {
    need(p && out_dev && psi_dev && n_peers >= 0 && n_peers <= 16, "pd_sharded_accumulate: bad argument");
    need(n_peers == 0 || (peer_slices && coef_host), "pd_sharded_accumulate: NULL peer list");
    for (int k = 0; k < n_peers; ++k)
      need(peer_slices[k] != nullptr, "pd_sharded_accumulate: NULL peer slice");
    p->eng.sharded_accumulate((pd::amp_t*)out_dev, (const pd::amp_t*)psi_dev, shift, n_peers,
                              (const pd::amp_t* const*)peer_slices, (const pd::cplx*)coef_host, stream);
  });
}
int pd_sharded_accumulate_range(pd_plan* p, void* stream, void* out_dev, const void* psi_dev, double shift,
                                int32_t n_peers, const void* const* peer_slices, const double* coef_host,
                                uint64_t n_amp) {
  return guarded_on(p, [&] {
    need(p && out_dev && psi_dev && n_peers >= 0 && n_peers <= 16 && n_amp > 0,
         "pd_sharded_accumulate_range: bad argument");
    need(n_peers == 0 || (peer_slices && coef_host), "pd_sharded_accumulate_range: NULL peer list");
    for (int k = 0; k < n_peers; ++k)
      need(peer_slices[k] != nullptr, "pd_sharded_accumulate_range: NULL peer slice");
    p->eng.sharded_accumulate_n((pd::amp_t*)out_dev, (const pd::amp_t*)psi_dev, shift, n_peers,
                                (const pd::amp_t* const*)peer_slices, (const pd::cplx*)coef_host, (size_t)n_amp,
                                stream);
  });
}
int pd_bench_hpsi(pd_plan* p, void* stream, double t, int32_t reps, const void* in_dev,
                  void* out_dev, double* ms_per_apply_host) {
  return guarded_on(p, [&] {
    need(p && in_dev && out_dev && ms_per_apply_host && reps > 0, "pd_bench_hpsi: bad argument");
    *ms_per_apply_host = p->eng.bench_apply((const pd::amp_t*)in_dev, (pd::amp_t*)out_dev, t, reps, stream);
  });
}
int pd_bench_dp5_steps(pd_plan* p, void* stream, double t0, double dt, int32_t steps, void* y_dev,
                       double* ms_per_step_host) {
  return guarded_on(p, [&] {
    need(p && y_dev && ms_per_step_host && steps > 0, "pd_bench_dp5_steps: bad argument");
    *ms_per_step_host = p->eng.bench_dp5((pd::amp_t*)y_dev, t0, dt, steps, stream);
  });
}
int pd_transfer_counters(int64_t* h2d_bytes, int64_t* d2h_bytes, int32_t reset) {
  return guarded([&] {
    long long a = 0, b = 0;
    PD_BACKEND::transfer_counters(&a, &b, reset != 0);
    if (h2d_bytes) *h2d_bytes = a;
    if (d2h_bytes) *d2h_bytes = b;
  });
}
int64_t pd_plan_launch_count(const pd_plan* p) { return p ? p->eng.launches : 0; }

}  // extern "C"
