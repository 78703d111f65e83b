// Evolution engine: DP5 / Krylov stepping, step controller, discrete adjoint.
//
// Host-side orchestration only -- every vector operation is a backend call (CUDA kernels
// in the product build; a plain-C++ stand-in under tests/emu for CPU tests of this logic).
//
// What it replaces (reference = /root/reference/pulser_diff, upstream = pyqtorch, unpinned):
//   * the solver loop behind backend.py:488-494 / 502-509   
//   * the autograd tape behind derivative.py:40,76           -- here a
//     discrete adjoint of the SAME accepted-step sequence, so gradients agree with the tape to
//     round-off, not merely to solver tolerance.
#pragma once
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <memory>

#include "pd_common.hpp"

namespace pd {

struct SlotInfo {      // one generator application inside the adjoint sweep
  double t_stage;      // time H was evaluated at
  double alpha;        // t_stage = t_step + alpha * h_step
  int interval;        // tsave interval the step belongs to
  int step;            // index into accepted-step list
};

struct AcceptedStep {
  double t, dt;
  int interval;
  int clipped;
};

struct Tape {
  std::vector<pd_step_record> records;   // every attempt
  std::vector<AcceptedStep> steps;       // accepted only, in order
  std::vector<double> tsave;
  int solver = PD_SOLVER_DP5_SE;
  pd_options opt;
  // Krylov: per interval and column, the Lanczos data needed by the adjoint
  struct KrylovSeg { int interval; double delta; double t_eval; };
  std::vector<KrylovSeg> ksegs;
  // identity of the device-side stage tape recorded by the small-register forward sweep (0 = none)
  uint64_t small_gen = 0;
  int unit_index = 0;   // position in a batch of parameter sets
  int n_acc_dev = -1, n_att_dev = -1;   // batch whose step list stayed on the device: its counts
};

template <class BK>
class Engine {
 public:
  using vec = amp_t*;
  BK bk;
  Program prog;
  Geometry geo{};
  Tableau tab;
  int64_t launches = 0;
  double* d_diag = nullptr;
  double* d_diag_parts = nullptr;
  size_t L = 0;  // complex elements per state block = batch * dim

  Engine(int nq, int batch, int kind, int device) : bk(device) {
    if (nq < 1 || batch < 1) throw Error(PD_ERR_INVALID, "n_qubits and batch must be >= 1");
    if (kind == PD_KET && nq > kMaxQubits) throw Error(PD_ERR_INVALID, "too many qubits");
    if (kind == PD_DENSITY && nq > kMaxSitesDensity)
      throw Error(PD_ERR_INVALID, "density plans support at most 16 qubits");
    if (kind != PD_KET && kind != PD_DENSITY) throw Error(PD_ERR_INVALID, "bad kind");
    prog.nq = nq;
    prog.kind = kind;
    prog.pair_u.assign((size_t)nq * nq, 0.0);
    geo.kind = kind;
    geo.nq = nq;
    geo.nbits = kind == PD_KET ? nq : 2 * nq;
    geo.dim = (size_t)1 << geo.nbits;
    geo.batch = batch;
    L = geo.dim * (size_t)batch;
    d_diag = (double*)bk.alloc(sizeof(double) * ((size_t)1 << nq));
    bk.zero(d_diag, sizeof(double) * ((size_t)1 << nq), nullptr);
    geo.diag = d_diag;
    if (kind == PD_KET && nq >= 16) {
      const size_t tiles = (size_t)1 << (nq - 12);
      d_diag_parts = (double*)bk.alloc(sizeof(double) * (4096 + 13 * tiles));
      bk.zero(d_diag_parts, sizeof(double) * (4096 + 13 * tiles), nullptr);
      geo.diag_parts = d_diag_parts;
    }
  }
  ~Engine() {
    for (auto& kv : bufs_) bk.free(kv.second);
    bk.free(d_diag);
    if (d_diag_parts) bk.free(d_diag_parts);
  }

  // ---- setup ----------------------------------------------------------------------------
  void set_interaction(const double* pair_u, void* stream) {
    std::copy(pair_u, pair_u + (size_t)prog.nq * prog.nq, prog.pair_u.begin());
    bk.build_diag(d_diag, prog.nq, prog.pair_u.data(), stream);
    if (d_diag_parts) bk.build_diag_parts(d_diag_parts, prog.nq, d_diag, stream);
    bk.sync(stream);
  }
  void set_terms(int n_samples, double dt, int n_det, const uint64_t* dm, const double* dv,
                 int n_amp, const uint64_t* am, const double* av) {
    if (n_samples < 2 || !(dt > 0)) throw Error(PD_ERR_INVALID, "need n_samples >= 2 and dt > 0");
    prog.n_samples = n_samples;
    prog.dt = dt;
    prog.det_masks.assign(dm, dm + n_det);
    prog.amp_masks.assign(am, am + n_amp);
    prog.det_values.assign(dv, dv + (size_t)n_det * n_samples);
    prog.amp_values.assign(av, av + (size_t)n_amp * n_samples * 2);
    ++prog.version;
  }
  void set_collapse(int n_ops, const double* ops) {
    if (prog.kind != PD_DENSITY) throw Error(PD_ERR_INVALID, "collapse operators need a density plan");
    prog.n_collapse = n_ops;
    prog.collapse.resize((size_t)n_ops * 4);
    for (size_t i = 0; i < (size_t)n_ops * 4; ++i) prog.collapse[i] = {ops[2 * i], ops[2 * i + 1]};
    prog.build_dsup();
  }

  // ---- one generator application -------------------------------------------------------
  // out = G(t) (sum_j w_j in_j);  comb (nullable) receives the combined input.
  void stage(vec out, vec comb, int n_in, const amp_t* const* ins, const double* w, double t,
             int mode, void* stream) {
    if (prog.kind == PD_KET) {
      SiteOps so;
      prog.site_ops_ket(t, mode, so);
      launches += bk.stage_ket(geo, out, comb, n_in, ins, w, so, scratch(stream), stream);
    } else {
      if (mode == 2) throw Error(PD_ERR_INVALID, "plain H apply is defined for ket plans only");
      SiteOpsDensity so;
      prog.site_ops_density(t, mode, so);
      launches += bk.stage_density(geo, out, comb, n_in, ins, w, so, scratch(stream), stream);
    }
  }
  void apply(vec out, const amp_t* in, double t, int mode, void* stream) {
    const amp_t* ins[1] = {in};
    double w[1] = {1.0};
    stage(out, nullptr, 1, ins, w, t, mode, stream);
  }

  // ---- forward ---------------------------------------------------------------------------
  void forward(int solver, const pd_options& opt, const amp_t* state0, const double* tsave, int n_t,
               amp_t* states, Tape* tape, void* stream) {
    if (n_t < 1) throw Error(PD_ERR_INVALID, "tsave must hold at least one time");
    for (int k = 1; k < n_t; ++k)
      if (tsave[k] < tsave[k - 1]) throw Error(PD_ERR_INVALID, "tsave must be sorted");
    bool me = solver == PD_SOLVER_DP5_ME;
    if (me != (prog.kind == PD_DENSITY))
      throw Error(PD_ERR_INVALID, "solver does not match the plan kind (DP5_ME <=> density)");
    if (tape) {
      tape->records.clear(); tape->steps.clear(); tape->ksegs.clear();
      tape->small_gen = 0;
      tape->tsave.assign(tsave, tsave + n_t);
      tape->solver = solver;
      tape->opt = opt;
      tape->opt.replay_dt = nullptr; tape->opt.replay_clipped = nullptr; tape->opt.n_replay = 0;
    }
    if (solver == PD_SOLVER_DP5_SE || solver == PD_SOLVER_DP5_ME)
      forward_dp5(opt, state0, tsave, n_t, states, tape, stream);
    else if (solver == PD_SOLVER_KRYLOV_SE)
      forward_krylov(opt, state0, tsave, n_t, states, tape, stream);
    else
      throw Error(PD_ERR_INVALID, "Solver not available.");
  }

  // ---- backward --------------------------------------------------------------------------
  void backward(Tape& tape, const amp_t* states, const amp_t* gstates, double* g_det, double* g_amp,
                double* g_pair, double* g_tsave, amp_t* g_state0, void* stream) {
    int n_t = (int)tape.tsave.size();
    int n_det = prog.n_det(), n_amp = prog.n_amp(), ns = prog.n_samples;
    if (g_det) std::fill(g_det, g_det + (size_t)n_det * ns, 0.0);
    if (g_amp) std::fill(g_amp, g_amp + (size_t)n_amp * ns * 2, 0.0);
    if (g_pair) std::fill(g_pair, g_pair + (size_t)prog.nq * prog.nq, 0.0);
    if (g_tsave) std::fill(g_tsave, g_tsave + n_t, 0.0);
    bool want_coef = g_det || g_amp || g_tsave;
    if (tape.solver == PD_SOLVER_KRYLOV_SE)
      backward_krylov(tape, states, gstates, g_det, g_amp, g_pair, g_tsave, g_state0, want_coef,
                      stream);
    else
      backward_dp5(tape, states, gstates, g_det, g_amp, g_pair, g_tsave, g_state0, want_coef,
                   stream);
  }

  // ---- batches of independent parameter sets (configs[2]) -------------------------------------
  // Same register, masks, time grid and options; unit u has its own coefficient tables
  // dv[u] ([n_det][n_samples]) / av[u] ([n_amp][n_samples] complex) and initial state.  One launch
  // evolves all units (one CTA per unit, small_ket*.cu); states: [U][n_t][batch][dim].
  void forward_units(const pd_options& o, int n_units, const amp_t* state0, const double* tsave, int n_t,
                     const double* dv, const double* av, amp_t* states, std::vector<Tape>* tapes,
                     uint64_t* gen_out, void* stream) {
    if (n_t < 1) throw Error(PD_ERR_INVALID, "tsave must hold at least one time");
    for (int k = 1; k < n_t; ++k)
      if (tsave[k] < tsave[k - 1]) throw Error(PD_ERR_INVALID, "tsave must be sorted");
    if (gen_out) *gen_out = 0;
    if (prog.kind != PD_KET) throw Error(PD_ERR_INVALID, "batches of parameter sets need a ket plan");
    if (!use_small() || !bk.small_units_supported(geo, prog)) {
      // units too large for one CTA each (or no cooperative kernels in this build): one after the
      // other through the single-problem path, each with its own coefficient tables
      if (BK::on_device(dv) || BK::on_device(av))
        throw Error(PD_ERR_INVALID, "device-resident coefficient tables need units that fit the one-launch kernels");
      size_t nd = (size_t)prog.n_det() * prog.n_samples, na = (size_t)prog.n_amp() * prog.n_samples * 2;
      if (tapes) tapes->assign(n_units, Tape{});
      for (int u = 0; u < n_units; ++u) {
        std::copy(dv + u * nd, dv + (u + 1) * nd, prog.det_values.begin());
        std::copy(av + u * na, av + (u + 1) * na, prog.amp_values.begin());
        ++prog.version;
        forward(PD_SOLVER_DP5_SE, o, state0 + (size_t)u * L, tsave, n_t, states + (size_t)u * n_t * L,
                tapes ? &(*tapes)[u] : nullptr, stream);
        if (tapes) (*tapes)[u].small_gen = 0;      // the device tape is overwritten by the next unit
      }
      return;
    }
    std::vector<std::vector<pd_step_record>> recs;
    uint64_t gen = 0;
    launches += bk.small_forward(geo, prog, tab, o, n_units, state0, dv, av, tsave, n_t, states, recs,
                                 tapes != nullptr, &gen, stream);
    if (gen_out) *gen_out = gen;
    if (tapes) {
      tapes->assign(n_units, Tape{});
      for (int u = 0; u < n_units; ++u) {
        Tape& t = (*tapes)[u];
        t.tsave.assign(tsave, tsave + n_t);
        t.solver = PD_SOLVER_DP5_SE;
        t.opt = o;
        t.opt.replay_dt = nullptr; t.opt.replay_clipped = nullptr; t.opt.n_replay = 0;
        t.small_gen = gen;
        t.unit_index = u;
        for (const auto& r : recs[u])
          if (r.accepted) t.steps.push_back({r.t, r.dt, r.interval, r.clipped});
        t.records = std::move(recs[u]);   // empty for batches: their step lists stay on the device
        if (n_units > 1) bk.small_unit_counts(gen, u, &t.n_acc_dev, &t.n_att_dev);
      }
    }
  }
  // g_det: [U][n_det][n_samples], g_amp: [U][n_amp][n_samples][2], g_state0: [U][batch][dim] (device)
  const amp_t* states_for_fallback_ = nullptr;   // set by the C ABI around backward_units
  void backward_units(std::vector<Tape>& tapes, const double* dv, const double* av, const amp_t* gstates,
                      double* g_det, double* g_amp, amp_t* g_state0, void* stream) {
    int n_units = (int)tapes.size();
    if (n_units == 0) return;
    int ns = prog.n_samples, n_det = prog.n_det(), n_amp = prog.n_amp();
    size_t nred = (size_t)n_det + 2 * (size_t)n_amp + 1;
    std::vector<std::vector<SkStepHost>> st(n_units);
    for (int u = 0; u < n_units; ++u)
      for (const auto& a : tapes[u].steps) st[u].push_back({a.t, a.dt, a.interval, a.clipped});
    amp_t* lam = (amp_t*)buf("lam_units", sizeof(amp_t) * L * (size_t)n_units);
    std::vector<std::vector<double>> sums;
    int nl = 0;
    bool want_coef = g_det || g_amp;
    if (tapes[0].small_gen != 0 && n_units > 1) {
      // step lists and gradient scatter on the device; only the sample gradients come back
      nl = bk.small_backward_units(geo, prog, tab, tapes[0].tsave, n_units, dv, av, tapes[0].small_gen, gstates,
                                   lam, g_det, g_amp, stream);
      if (nl > 0) {
        launches += nl;
        if (g_state0) bk.d2d(g_state0, lam, sizeof(amp_t) * L * (size_t)n_units, stream);
        bk.sync(stream);
        return;
      }
    }
    if (BK::on_device(dv) || BK::on_device(av) || BK::on_device(g_det) || BK::on_device(g_amp))
      throw Error(PD_ERR_INVALID, "device-resident coefficient tables / gradients need a batch (> 1 unit) whose "
                                  "device-side tape is still current");
    if (tapes[0].small_gen != 0 && !tapes[0].steps.empty())
      nl = bk.small_backward(geo, prog, tab, tapes[0].tsave, n_units, dv, av, st, tapes[0].small_gen, gstates,
                             want_coef, nullptr, lam, sums, stream);
    if (nl == 0 && tapes[0].small_gen != 0 && n_units > 1)
      throw Error(PD_ERR_STATE, "the device-side tape of this batch was overwritten by a later evolution on the "
                                "same plan; run the batch forward again before its backward");
    if (nl == 0) {
      // no device tape (large units, another evolution ran since, or a build without the cooperative
      // kernels): stage-by-stage adjoint per unit, recomputing from the saved states
      if (!states_for_fallback_)
        throw Error(PD_ERR_STATE, "backward_units: the forward states are needed for the per-unit adjoint");
      size_t nd = (size_t)n_det * ns, na = (size_t)n_amp * ns * 2;
      int n_t = (int)tapes[0].tsave.size();
      for (int u = 0; u < n_units; ++u) {
        std::copy(dv + u * nd, dv + (u + 1) * nd, prog.det_values.begin());
        std::copy(av + u * na, av + (u + 1) * na, prog.amp_values.begin());
        ++prog.version;
        backward(tapes[u], states_for_fallback_ + (size_t)u * n_t * L,
                 gstates ? gstates + (size_t)u * n_t * L : nullptr, g_det ? g_det + u * nd : nullptr,
                 g_amp ? g_amp + u * na : nullptr, nullptr, nullptr,
                 g_state0 ? g_state0 + (size_t)u * L : nullptr, stream);
      }
      return;
    }
    launches += nl;
    if (g_state0) bk.d2d(g_state0, lam, sizeof(amp_t) * L * (size_t)n_units, stream);
    if (g_det) std::fill(g_det, g_det + (size_t)n_units * n_det * ns, 0.0);
    if (g_amp) std::fill(g_amp, g_amp + (size_t)n_units * n_amp * ns * 2, 0.0);
    if (want_coef)
      for (int u = 0; u < n_units; ++u)
        for (size_t gi = 0; gi < tapes[u].steps.size(); ++gi) {
          const AcceptedStep& s = tapes[u].steps[gi];
          for (int i = 0; i < 6; ++i) {
            double alpha = i == 0 ? 0.0 : tab.alpha[i - 1];
            distribute_terms(s.t + s.dt * alpha, &sums[u][(gi * 6 + i) * nred],
                             g_det ? g_det + (size_t)u * n_det * ns : nullptr,
                             g_amp ? g_amp + (size_t)u * n_amp * ns * 2 : nullptr);
          }
        }
    bk.sync(stream);
  }

  // ---- measurement hooks ------------------------------------------------------------------
  double bench_apply(const amp_t* in, amp_t* out, double t, int reps, void* stream) {
    apply(out, in, t, 2, stream);
    bk.sync(stream);
    bk.timer_start(stream);
    for (int r = 0; r < reps; ++r) apply(out, in, t, 2, stream);
    return bk.timer_stop_ms(stream) / std::max(1, reps);
  }
  double bench_dp5(amp_t* y_io, double t0, double dt, int steps, void* stream) {
    vec y = vbuf("y"), ynew = vbuf("ynew");
    vec k[7];
    for (int i = 0; i < 7; ++i) k[i] = vbuf("k" + std::to_string(i));
    double* d_err = (double*)buf("norm_out", sizeof(double) * geo.batch);
    bk.d2d(y, y_io, sizeof(amp_t) * L, stream);
    double t = t0;
    apply(k[0], y, t, 0, stream);
    auto one = [&]() {
      dp5_step_with_error(t, dt, y, k, ynew, 1e-8, 1e-6, d_err, stream);
      t += dt;
      std::swap(y, ynew);
      std::swap(k[0], k[6]);
    };
    one();
    bk.sync(stream);
    bk.timer_start(stream);
    for (int s = 0; s < steps; ++s) one();
    double ms = bk.timer_stop_ms(stream) / std::max(1, steps);
    bk.d2d(y_io, y, sizeof(amp_t) * L, stream);
    bk.sync(stream);
    return ms;
  }

  // ---- diagonal expectation --------------------------------------------------------------
  void expect_diag(const amp_t* states, int n_t, const double* obs, double* out_host, void* stream) {
    cplx* d_out = (cplx*)buf("expect", sizeof(cplx) * (size_t)n_t);
    launches += bk.expect_diag(geo, states, n_t, obs, d_out, reduce_scratch(), stream);
    bk.d2h(out_host, d_out, sizeof(cplx) * (size_t)n_t, stream);
    bk.sync(stream);
  }

  // ---- vector-Jacobian product of ONE generator application ---------------------------------
  // k = G(t) y with cotangent kbar:  grad_y = G(t)^dagger kbar; the coefficient-sample gradients
  // are ADDED into g_det / g_amp, g_pair is overwritten; returns dL/dt.  Same reductions as one
  // stage of the DP5 adjoint (adjoint_step), exposed so host-side integrators (the sharded
  // register, user-written steppers) are differentiable too.
  double rhs_vjp(double t, const amp_t* y, const amp_t* kbar, amp_t* grad_y, double* g_det,
                 double* g_amp, double* g_pair, bool defer_pair, void* stream) {
    if (grad_y) apply(grad_y, kbar, t, 1, stream);
    int cs = corr_stride();
    cplx* d_corr = (cplx*)buf("corr", sizeof(cplx) * (size_t)cs);
    double* d_wacc = nullptr;
    size_t wbytes = sizeof(double) * ((size_t)1 << geo.nq);
    if (defer_pair) {
      // per-amplitude weights keep accumulating inside the plan; pair_gradient_flush() reduces them
      d_wacc = (double*)buf("wacc_vjp", wbytes);
      if (!vjp_wacc_live_) { bk.zero(d_wacc, wbytes, stream); vjp_wacc_live_ = true; }
    } else if (g_pair) {
      d_wacc = (double*)buf("wacc", wbytes);
      bk.zero(d_wacc, wbytes, stream);
    }
    const amp_t* yi[1] = {y};
    double yw[1] = {1.0};
    launches += bk.corr_combo(geo, d_corr, d_wacc, 1.0, kbar, 1, yi, yw, vbuf("ystage"),
                              reduce_scratch(), stream);
    std::vector<cplx> h_corr(cs);
    bk.d2h(h_corr.data(), d_corr, sizeof(cplx) * (size_t)cs, stream);
    if (g_pair && !defer_pair) {
      double* d_pair = (double*)buf("pair_out", sizeof(double) * (size_t)prog.nq * prog.nq);
      launches += bk.pair_reduce(geo, d_pair, d_wacc, stream);
      bk.d2h(g_pair, d_pair, sizeof(double) * (size_t)prog.nq * prog.nq, stream);
    }
    bk.sync(stream);
    return distribute(t, h_corr.data(), g_det, g_amp);
  }
  // dL/dU_ij of every deferred rhs_vjp since the last flush; clears the accumulator
  void pair_gradient_flush(double* g_pair, void* stream) {
    size_t n2 = (size_t)prog.nq * prog.nq;
    if (!vjp_wacc_live_) { std::fill(g_pair, g_pair + n2, 0.0); return; }
    double* d_wacc = (double*)buf("wacc_vjp", sizeof(double) * ((size_t)1 << geo.nq));
    double* d_pair = (double*)buf("pair_out", sizeof(double) * n2);
    launches += bk.pair_reduce(geo, d_pair, d_wacc, stream);
    bk.d2h(g_pair, d_pair, sizeof(double) * n2, stream);
    bk.sync(stream);
    vjp_wacc_live_ = false;
  }
  bool vjp_wacc_live_ = false;
  size_t n_backward_ = 0;   // DP5 adjoint sweeps this plan has run (slope-cache policy)

  // ---- building blocks of a host-driven DP5 step (the sharded register) --------------------
  // out = sum_j w_j in_j (one pass)
  void lincomb(amp_t* out, int n_in, const amp_t* const* ins, const double* w, void* stream) {
    launches += bk.lincomb(geo, out, n_in, ins, w, stream);
  }
  // per-column sum over this plan's amplitudes of |sum_j ew_j k_j / (atol + rtol max(|y0|,|y1|))|^2
  void error_sumsq(const amp_t* const* k, const double* ew, const amp_t* y0, const amp_t* y1,
                   double atol, double rtol, double* out_host, void* stream) {
    double* d_err = (double*)buf("norm_out", sizeof(double) * geo.batch);
    launches += bk.err_sumsq(geo, d_err, k, ew, y0, y1, atol, rtol, reduce_scratch(), stream);
    bk.d2h(out_host, d_err, sizeof(double) * geo.batch, stream);
    bk.sync(stream);
  }

  // ---- sharded register: flips of the qubits that index the rank -----------
  // out += shift*psi + sum_k coef_k * peers[k]; peers[k] may be peer-mapped device memory.
  void sharded_accumulate(amp_t* out, const amp_t* psi, double shift, int n_peers,
                          const amp_t* const* peers, const cplx* coef, void* stream) {
    launches += bk.sharded_accumulate(geo, out, psi, shift, n_peers, peers, coef, stream);
  }
  void sharded_accumulate_n(amp_t* out, const amp_t* psi, double shift, int n_peers, const amp_t* const* peers,
                            const cplx* coef, size_t n_amp, void* stream) {
    if (n_amp > geo.dim * (size_t)geo.batch) throw Error(PD_ERR_INVALID, "range longer than the slice");
    Geometry g1 = geo;
    g1.batch = 1;
    g1.dim = n_amp;
    launches += bk.sharded_accumulate(g1, out, psi, shift, n_peers, peers, coef, stream);
  }

 private:
  std::vector<std::pair<std::string, void*>> bufs_;
  std::vector<std::pair<std::string, size_t>> buf_sizes_;

  void* buf(const std::string& name, size_t bytes) {
    for (size_t i = 0; i < bufs_.size(); ++i)
      if (bufs_[i].first == name) {
        if (buf_sizes_[i].second >= bytes) return bufs_[i].second;
        bk.free(bufs_[i].second);
        bufs_[i].second = bk.alloc(bytes);
        buf_sizes_[i].second = bytes;
        return bufs_[i].second;
      }
    void* p = bk.alloc(bytes);
    bufs_.push_back({name, p});
    buf_sizes_.push_back({name, bytes});
    return p;
  }
  size_t buf_bytes(const std::string& name) const {
    for (size_t i = 0; i < bufs_.size(); ++i)
      if (bufs_[i].first == name) return buf_sizes_[i].second;
    return 0;
  }
  void drop(const std::string& name) {
    for (size_t i = 0; i < bufs_.size(); ++i)
      if (bufs_[i].first == name) {
        bk.free(bufs_[i].second);
        bufs_.erase(bufs_.begin() + i);
        buf_sizes_.erase(buf_sizes_.begin() + i);
        return;
      }
  }
  vec vbuf(const std::string& name) { return (vec)buf(name, sizeof(amp_t) * L); }
  vec scratch(void*) { return vbuf("scratch"); }
  double* reduce_scratch() { return (double*)buf("reduce", bk.reduce_scratch_bytes(geo)); }

  double hairer_from_sumsq(const double* sumsq) const {
    double m = 0.0;
    for (int b = 0; b < geo.batch; ++b) m = std::max(m, std::sqrt(sumsq[b] / (double)geo.dim));
    return m;
  }
  double scaled_norm(const amp_t* x, const amp_t* xsub, const amp_t* ref, double atol, double rtol,
                     void* stream) {
    double* d = (double*)buf("norm_out", sizeof(double) * geo.batch);
    std::vector<double> h(geo.batch);
    launches += bk.scaled_sumsq(geo, d, x, xsub, ref, atol, rtol, reduce_scratch(), stream);
    bk.d2h(h.data(), d, sizeof(double) * geo.batch, stream);
    bk.sync(stream);
    return hairer_from_sumsq(h.data());
  }

  // Hairer's initial step heuristic
  double init_tstep(double t0, const amp_t* y0, const amp_t* f0, const pd_options& o, void* stream) {
    double d0 = scaled_norm(y0, nullptr, y0, o.atol, o.rtol, stream);
    double d1 = scaled_norm(f0, nullptr, y0, o.atol, o.rtol, stream);
    double h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
    vec y1 = vbuf("ynew"), f1 = vbuf("k1");
    const amp_t* ins[2] = {y0, f0};
    double w[2] = {1.0, h0};
    stage(f1, y1, 2, ins, w, t0 + h0, 0, stream);
    double d2 = scaled_norm(f1, f0, y0, o.atol, o.rtol, stream) / h0;
    double h1 = (d1 <= 1e-15 && d2 <= 1e-15) ? std::max(1e-6, h0 * 1e-3)
                                             : std::pow(0.01 / std::max(d1, d2), 1.0 / 6.0);
    return std::min(100 * h0, h1);
  }
  static double update_tstep(double dt, double error, const pd_options& o) {
    if (error == 0.0) return dt * o.max_factor;
    double fac = o.safety_factor * std::pow(error, -1.0 / 5.0);
    if (error <= 1.0) return dt * std::max(1.0, std::min(o.max_factor, fac));
    return dt * std::min(0.9, std::max(o.min_factor, fac));
  }

  // stages 2..(last) of one DP5 step from (t, y, k[0]); fills k[1..last-1]; ynew if last == 7.
  // ysave (nullable): ysave[i] receives the input of stage i+1, Y_i = y + dt sum_j beta_ij k_j, i = 1..5 --
  // the launch that applies the generator forms and writes it anyway; the adjoint sweep reads it back for the
  // site correlations instead of re-forming it from the slopes.
  void dp5_stages(double t, double dt, const amp_t* y, vec* k, vec ynew, int last, void* stream,
                  vec* ysave = nullptr) {
    for (int i = 1; i < last; ++i) {
      const amp_t* ins[8];
      double w[8];
      int n = 0;
      ins[n] = y; w[n++] = 1.0;
      for (int j = 0; j < i; ++j) {
        double b = tab.beta[i - 1][j];
        if (b != 0.0) { ins[n] = k[j]; w[n++] = dt * b; }
      }
      stage(k[i], i == 6 ? ynew : (ysave ? ysave[i] : nullptr), n, ins, w, t + dt * tab.alpha[i - 1], 0, stream);
    }
  }

  // stages 2..7, y_{n+1} and the per-column error sums of one step; fused tiled path if the
  // backend offers one for this shape, stage by stage otherwise
  void dp5_step_with_error(double t, double dt, const amp_t* y, vec* k, vec ynew, double atol,
                           double rtol, double* d_err, void* stream) {
    double ew[7];
    for (int j = 0; j < 7; ++j) ew[j] = dt * (tab.b5[j] - tab.b4[j]);
    if (prog.kind == PD_KET) {
      std::vector<SiteOps> so(7);
      for (int i = 1; i < 7; ++i) prog.site_ops_ket(t + dt * tab.alpha[i - 1], 0, so[i]);
      int nl = bk.dp5_step_ket(geo, y, k, ynew, so.data(), tab, ew, dt, atol, rtol, vbuf("scratch"),
                               vbuf("scratch2"), reduce_scratch(), d_err, stream);
      if (nl > 0) { launches += nl; return; }
    }
    dp5_stages(t, dt, y, k, ynew, 7, stream);
    launches += bk.err_sumsq(geo, d_err, (const amp_t* const*)k, ew, y, ynew, atol, rtol,
                             reduce_scratch(), stream);
  }

  void forward_dp5(const pd_options& o, const amp_t* state0, const double* tsave, int n_t,
                   amp_t* states, Tape* tape, void* stream) {
    vec y = vbuf("y"), ynew = vbuf("ynew");
    vec k[7];
    for (int i = 0; i < 7; ++i) k[i] = vbuf("k" + std::to_string(i));
    if (use_small()) {
      // whole evolution in one cooperative kernel (small_ket*.cu), initial slope and step included;
      // the attempt log becomes the tape
      std::vector<std::vector<pd_step_record>> recs;
      uint64_t gen = 0;
      launches += bk.small_forward(geo, prog, tab, o, 1, state0, nullptr, nullptr, tsave, n_t, states, recs,
                                   tape != nullptr, &gen, stream);
      if (tape) {
        tape->small_gen = gen;
        for (const auto& r : recs[0])
          if (r.accepted) tape->steps.push_back({r.t, r.dt, r.interval, r.clipped});
        tape->records = std::move(recs[0]);
      }
      return;
    }
    bk.d2d(y, state0, sizeof(amp_t) * L, stream);
    double t = tsave[0];
    apply(k[0], y, t, 0, stream);
    bool replay = o.n_replay > 0;
    double dt = replay ? 0.0 : init_tstep(t, y, k[0], o, stream);
    double error = 1.0;
    int64_t pos = 0;
    double* d_err = (double*)buf("norm_out", sizeof(double) * geo.batch);
    std::vector<double> h_err(geo.batch);
    for (int kk = 0; kk < n_t; ++kk) {
      double t_next = tsave[kk];
      double cache_dt = dt, cache_err = error;
      int64_t steps = 0;
      while (t < t_next) {
        bool clipped;
        if (!replay) {
          dt = update_tstep(dt, error, o);
          clipped = t + dt >= t_next;
        } else {
          if (pos >= o.n_replay) throw Error(PD_ERR_INVALID, "replay sequence too short");
          dt = o.replay_dt[pos];
          clipped = o.replay_clipped[pos] != 0;
          ++pos;
        }
        if (clipped) { cache_dt = dt; cache_err = error; dt = t_next - t; }
        dp5_step_with_error(t, dt, y, k, ynew, o.atol, o.rtol, d_err, stream);
        bk.d2h(h_err.data(), d_err, sizeof(double) * geo.batch, stream);
        bk.sync(stream);
        error = hairer_from_sumsq(h_err.data());
        bool accepted = replay ? true : error <= 1.0;   // a replayed sequence holds accepted steps only
        if (!(error == error)) throw Error(PD_ERR_STATE, "non-finite error norm in DP5 step");
        if (tape) {
          tape->records.push_back({t, dt, error, accepted ? 1 : 0, clipped ? 1 : 0, kk, 0});
          if (accepted) tape->steps.push_back({t, dt, kk, clipped ? 1 : 0});
        }
        if (accepted) {
          t = clipped ? t_next : t + dt;
          std::swap(y, ynew);
          std::swap(k[0], k[6]);  // FSAL
        }
        if (++steps >= o.max_steps) throw Error(PD_ERR_MAX_STEPS, "max_steps reached");
      }
      dt = cache_dt; error = cache_err;
      bk.d2d(states + (size_t)kk * L, y, sizeof(amp_t) * L, stream);
    }
    bk.sync(stream);
  }

  // ---- DP5 adjoint -------------------------------------------------------------------------
  int corr_stride() const { return geo.nq * (prog.kind == PD_KET ? 4 : 16); }

  void backward_dp5(Tape& tape, const amp_t* states, const amp_t* gstates, double* g_det,
                    double* g_amp, double* g_pair, double* g_tsave, amp_t* g_state0, bool want_coef,
                    void* stream) {
    int n_t = (int)tape.tsave.size();
    size_t n_steps = tape.steps.size();
    vec lam = vbuf("lam");
    if (use_small() &&
        backward_dp5_small(tape, states, gstates, g_det, g_amp, g_pair, g_tsave, g_state0, want_coef, stream))
      return;
    // PD_TIMING=1: host-side phase times of this sweep on stderr
    static const bool timing = std::getenv("PD_TIMING") != nullptr;
    auto now = [&] { if (timing) bk.sync(stream); return std::chrono::steady_clock::now(); };
    auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto t_begin = now();
    double ms_alloc = 0, ms_recompute = 0, ms_sweep = 0;
    if (gstates) bk.d2d(lam, gstates + (size_t)(n_t - 1) * L, sizeof(amp_t) * L, stream);
    else bk.zero(lam, sizeof(amp_t) * L, stream);
    vec k[6], yb[6], ys[6];
    for (int i = 0; i < 6; ++i) { k[i] = vbuf("k" + std::to_string(i)); yb[i] = vbuf("yb" + std::to_string(i)); }
    ys[0] = nullptr;
    for (int i = 1; i < 6; ++i) ys[i] = vbuf("ys" + std::to_string(i));
    vec kbar = vbuf("kbar"), ystage = vbuf("ystage");
    // per-slot reduction results
    size_t n_slots = n_steps * 6;
    int cs = corr_stride();
    cplx* d_corr = want_coef ? (cplx*)buf("corr", sizeof(cplx) * std::max<size_t>(1, n_slots * cs)) : nullptr;
    double* d_hdot = g_tsave ? (double*)buf("hdot", sizeof(double) * std::max<size_t>(1, n_slots)) : nullptr;
    if (d_hdot) bk.zero(d_hdot, sizeof(double) * std::max<size_t>(1, n_slots), stream);
    double* d_wacc = nullptr;
    if (g_pair) {
      d_wacc = (double*)buf("wacc", sizeof(double) * ((size_t)1 << geo.nq));
      bk.zero(d_wacc, sizeof(double) * ((size_t)1 << geo.nq), stream);
    }
    std::vector<SlotInfo> slots(n_slots);

    // segment buffer for recomputed step-start states
    size_t vec_bytes = sizeof(amp_t) * L;
    // a third of the memory that is free or already held by this plan's segment buffers (a sweep that grew
    // them must not shrink its own budget on the next call)
    const size_t held = buf_bytes("seg") + buf_bytes("segk");
    size_t cap = std::max<size_t>(1, (bk.segment_budget_bytes() + held / 3) / vec_bytes);
    auto t_setup = now();

    size_t hi = n_steps;  // steps [lo, hi) belong to the interval being processed
    for (int kk = n_t - 1; kk >= 1; --kk) {
      size_t lo = hi;
      while (lo > 0 && tape.steps[lo - 1].interval == kk) --lo;
      size_t ns = hi - lo;
      if (ns > 0) {
        const amp_t* y_start = states + (size_t)(kk - 1) * L;
        // process the interval's steps in chunks of <= cap, last chunk first
        size_t done_hi = ns;
        while (done_hi > 0) {
          size_t c_lo = done_hi > cap ? done_hi - cap : 0;
          size_t cn = done_hi - c_lo;
          // y at the start of local step c_lo: recompute from the interval start
          auto t0 = now();
          vec seg = (vec)buf("seg", vec_bytes * cn);
          vec ycur = vbuf("y"), ynext = vbuf("ynew");
          const amp_t* ysrc = y_start;
          for (size_t s = 0; s < c_lo; ++s) {
            const AcceptedStep& st = tape.steps[lo + s];
            advance(ysrc, ynext, st, k, stream);
            std::swap(ycur, ynext);
            ysrc = ycur;
          }
          bk.d2d(seg, ysrc, vec_bytes, stream);
          // Spare segment budget keeps the slopes and stage inputs of the recomputation pass (6 + 5 vectors per
          // step) for as many steps as fit, so the sweep does not recompute them.  Allocating that cache costs
          // more than one sweep saves (cudaMalloc ~50 ms/GiB against ~3 ms/GiB of recomputation), so
          // it is only set up once a plan is differentiated repeatedly (optimisation loops).
          constexpr size_t kPerStep = 11;
          const size_t nk = (n_backward_ >= 1 && cap > cn) ? std::min(cn - 1, (cap - cn) / kPerStep) : 0;
          vec segk = nk ? (vec)buf("segk", vec_bytes * kPerStep * nk) : nullptr;
          auto cache_ptrs = [&](size_t s, vec* ks, vec* yss) {
            for (int i = 0; i < 6; ++i) ks[i] = s < nk ? segk + (s * kPerStep + i) * L : k[i];
            yss[0] = nullptr;
            for (int i = 1; i < 6; ++i) yss[i] = s < nk ? segk + (s * kPerStep + 5 + i) * L : ys[i];
          };
          auto t1 = now();
          ms_alloc += ms(t0, t1);
          for (size_t s = 0; s + 1 < cn; ++s) {
            const AcceptedStep& st = tape.steps[lo + c_lo + s];
            vec ks[6], yss[6];
            cache_ptrs(s, ks, yss);
            advance(seg + s * L, seg + (s + 1) * L, st, ks, stream, s < nk ? yss : nullptr);
          }
          auto t2 = now();
          ms_recompute += ms(t1, t2);
          for (size_t s = cn; s-- > 0;) {
            size_t gi = lo + c_lo + s;
            const bool cached = s < nk;
            vec ks[6], yss[6];
            cache_ptrs(s, ks, yss);
            adjoint_step(tape.steps[gi], (int)gi, seg + s * L, lam, ks, yss, yb, kbar, ystage, d_corr,
                         d_hdot, d_wacc, slots, want_coef, stream, cached);
          }
          ms_sweep += ms(t2, now());
          done_hi = c_lo;
        }
      }
      hi = lo;
      if (gstates) {
        const amp_t* ins[2] = {lam, gstates + (size_t)(kk - 1) * L};
        double w[2] = {1.0, 1.0};
        launches += bk.lincomb(geo, lam, 2, ins, w, stream);
      }
    }
    if (g_state0) bk.d2d(g_state0, lam, vec_bytes, stream);
    ++n_backward_;
    if (timing)
      std::fprintf(stderr, "[pd] backward_dp5: setup %.1f ms, segment alloc %.1f ms, recompute %.1f ms, sweep %.1f ms "
                   "(%zu steps, cap %zu vectors)\n", ms(t_begin, t_setup), ms_alloc, ms_recompute, ms_sweep, n_steps, cap);

    // ---- host post-processing of the per-slot reductions ----
    if (want_coef && n_slots > 0) {
      std::vector<cplx> h_corr(n_slots * cs);
      bk.d2h(h_corr.data(), d_corr, sizeof(cplx) * n_slots * cs, stream);
      std::vector<double> h_hdot;
      if (d_hdot) { h_hdot.resize(n_slots); bk.d2h(h_hdot.data(), d_hdot, sizeof(double) * n_slots, stream); }
      bk.sync(stream);
      std::vector<double> tbar_interval(n_t, 0.0), hbar_interval(n_t, 0.0);
      for (size_t si = 0; si < n_slots; ++si) {
        const SlotInfo& sl = slots[si];
        double tbar = distribute(sl.t_stage, &h_corr[si * cs], g_det, g_amp);
        tbar_interval[sl.interval] += tbar;
        const AcceptedStep& st = tape.steps[sl.step];
        if (st.clipped) hbar_interval[sl.interval] += sl.alpha * tbar + (d_hdot ? h_hdot[si] / st.dt : 0.0);
      }
      if (g_tsave)
        for (int kk = 1; kk < n_t; ++kk) {
          g_tsave[kk] += hbar_interval[kk];
          g_tsave[kk - 1] += tbar_interval[kk] - hbar_interval[kk];
        }
    }
    if (g_pair) {
      double* d_pair = (double*)buf("pair_out", sizeof(double) * (size_t)prog.nq * prog.nq);
      launches += bk.pair_reduce(geo, d_pair, d_wacc, stream);
      bk.d2h(g_pair, d_pair, sizeof(double) * (size_t)prog.nq * prog.nq, stream);
    }
    bk.sync(stream);
  }

  // ---- small-register family (one cooperative kernel per sweep) -------------------------------
  bool use_small() {
    if (prog.kind != PD_KET || !(bk.path == 0 || bk.path == 3)) return false;
    bool ok = bk.small_supported(geo, prog);
    if (bk.path == 3 && !ok) throw Error(PD_ERR_INVALID, "path 3 (small-register kernels) does not fit this plan");
    return ok;
  }
  bool backward_dp5_small(Tape& tape, const amp_t* states, const amp_t* gstates, double* g_det,
                          double* g_amp, double* g_pair, double* g_tsave, amp_t* g_state0,
                          bool want_coef, void* stream) {
    int n_t = (int)tape.tsave.size();
    int n_steps = (int)tape.steps.size();
    std::vector<std::vector<SkStepHost>> st(1);
    for (const auto& a : tape.steps) st[0].push_back({a.t, a.dt, a.interval, a.clipped});
    vec lam = vbuf("lam");
    double* d_wacc = nullptr;
    if (g_pair) {
      d_wacc = (double*)buf("wacc", sizeof(double) * ((size_t)1 << geo.nq));
      bk.zero(d_wacc, sizeof(double) * ((size_t)1 << geo.nq), stream);
    }
    std::vector<std::vector<double>> sums_u;
    int nl = bk.small_backward(geo, prog, tab, tape.tsave, 1, nullptr, nullptr, st, tape.small_gen, gstates,
                               want_coef, d_wacc, lam, sums_u, stream);
    if (nl == 0) return false;
    launches += nl;
    if (g_state0) bk.d2d(g_state0, lam, sizeof(amp_t) * L, stream);
    if (want_coef && n_steps > 0) {
      const std::vector<double>& sums = sums_u[0];
      size_t nred = (size_t)prog.n_det() + 2 * (size_t)prog.n_amp() + 1;
      std::vector<double> tbar_interval(n_t, 0.0), hbar_interval(n_t, 0.0);
      for (int gi = 0; gi < n_steps; ++gi) {
        const AcceptedStep& st = tape.steps[gi];
        for (int i = 0; i < 6; ++i) {
          double alpha = i == 0 ? 0.0 : tab.alpha[i - 1];
          double ts = st.t + st.dt * alpha;
          const double* sm = &sums[((size_t)gi * 6 + i) * nred];
          double tbar = distribute_terms(ts, sm, g_det, g_amp);
          tbar_interval[st.interval] += tbar;
          if (st.clipped) hbar_interval[st.interval] += alpha * tbar + (g_tsave ? sm[nred - 1] / st.dt : 0.0);
        }
      }
      if (g_tsave)
        for (int kk = 1; kk < n_t; ++kk) {
          g_tsave[kk] += hbar_interval[kk];
          g_tsave[kk - 1] += tbar_interval[kk] - hbar_interval[kk];
        }
    }
    if (g_pair) {
      double* d_pair = (double*)buf("pair_out", sizeof(double) * (size_t)prog.nq * prog.nq);
      launches += bk.pair_reduce(geo, d_pair, d_wacc, stream);
      bk.d2h(g_pair, d_pair, sizeof(double) * (size_t)prog.nq * prog.nq, stream);
    }
    bk.sync(stream);
    return true;
  }
  // Per-term sums of one slot (det: sum_q gd_q, amp: sum_q ga_q / gb_q) -> sample / time gradients.
  double distribute_terms(double ts, const double* sums, double* g_det, double* g_amp) const {
    int ns = prog.n_samples, n_det = prog.n_det();
    if (ns < 2) return 0.0;
    Interp ix = interp_index(ts, prog.dt, ns);
    double x = (ts - ix.i1 * prog.dt) / prog.dt;
    double tbar = 0.0;
    for (int kdx = 0; kdx < n_det; ++kdx) {
      double g = 2.0 * sums[kdx];
      const double* v = &prog.det_values[(size_t)kdx * ns];
      if (g_det) {
        g_det[(size_t)kdx * ns + ix.i1] += g * (1.0 - x);
        g_det[(size_t)kdx * ns + ix.i2] += g * x;
      }
      tbar += g * (v[ix.i2] - v[ix.i1]) / prog.dt;
    }
    for (int kdx = 0; kdx < prog.n_amp(); ++kdx) {
      double gre = sums[n_det + 2 * kdx], gim = sums[n_det + 2 * kdx + 1];
      const double* v = &prog.amp_values[(size_t)kdx * ns * 2];
      if (g_amp) {
        double* o1 = &g_amp[((size_t)kdx * ns + ix.i1) * 2];
        double* o2 = &g_amp[((size_t)kdx * ns + ix.i2) * 2];
        o1[0] += gre * (1.0 - x); o1[1] += gim * (1.0 - x);
        o2[0] += gre * x;         o2[1] += gim * x;
      }
      tbar += gre * (v[2 * ix.i2] - v[2 * ix.i1]) / prog.dt +
              gim * (v[2 * ix.i2 + 1] - v[2 * ix.i1 + 1]) / prog.dt;
    }
    return tbar;
  }

  // y_out = DP5 step from y_in (no error estimate); leaves k[0..5] filled
  void advance(const amp_t* y_in, vec y_out, const AcceptedStep& st, vec* k, void* stream, vec* ysave = nullptr) {
    apply(k[0], y_in, st.t, 0, stream);
    dp5_stages(st.t, st.dt, y_in, k, nullptr, 6, stream, ysave);
    const amp_t* ins[8];
    double w[8];
    int n = 0;
    ins[n] = y_in; w[n++] = 1.0;
    for (int j = 0; j < 6; ++j)
      if (tab.b5[j] != 0.0) { ins[n] = k[j]; w[n++] = st.dt * tab.b5[j]; }
    launches += bk.lincomb(geo, y_out, n, ins, w, stream);
  }

  // ys[1..5]: the stage inputs Y_i of this step (filled here unless the recomputation pass cached them with
  // the slopes, have_k)
  void adjoint_step(const AcceptedStep& st, int step_index, const amp_t* y_n, vec lam, vec* k, vec* ys,
                    vec* yb, vec kbar, vec ystage, cplx* d_corr, double* d_hdot, double* d_wacc,
                    std::vector<SlotInfo>& slots, bool want_coef, void* stream, bool have_k = false) {
    double t = st.t, h = st.dt;
    if (!have_k) {
      apply(k[0], y_n, t, 0, stream);
      // the sweep reads the stage inputs Y_1..Y_5 and, for a clipped step whose end time is differentiated, the
      // slopes (re_dot below): otherwise the sixth slope is never used and its application is skipped --
      // Y_5 is then formed by a plain combination
      const bool need_k5 = d_hdot && st.clipped;
      dp5_stages(t, h, y_n, k, nullptr, need_k5 ? 6 : 5, stream, ys);
      if (!need_k5) {
        const amp_t* ins[8];
        double w[8];
        int n = 0;
        ins[n] = y_n; w[n++] = 1.0;
        for (int j = 0; j < 5; ++j) {
          double b = tab.beta[4][j];
          if (b != 0.0) { ins[n] = k[j]; w[n++] = h * b; }
        }
        launches += bk.lincomb(geo, ys[5], n, ins, w, stream);
      }
    }
    int cs = corr_stride();
    for (int i = 5; i >= 0; --i) {
      const amp_t* ins[8];
      double w[8];
      int n = 0;
      if (tab.b5[i] != 0.0) { ins[n] = lam; w[n++] = h * tab.b5[i]; }
      for (int j = i + 1; j < 6; ++j) {
        double b = tab.beta[j - 1][i];
        if (b != 0.0) { ins[n] = yb[j]; w[n++] = h * b; }
      }
      double alpha = i == 0 ? 0.0 : tab.alpha[i - 1];
      double ts = t + h * alpha;
      stage(yb[i], kbar, n, ins, w, ts, 1, stream);
      size_t slot = (size_t)step_index * 6 + i;
      slots[slot] = {ts, alpha, st.interval, step_index};
      if (want_coef || d_wacc) {
        // stage input Y_i = y_n + h sum_j beta_ij k_j as the forward stage launch wrote it (Y_0 = y_n)
        const amp_t* yi[1] = {i == 0 ? y_n : ys[i]};
        const double yw[1] = {1.0};
        launches += bk.corr_combo(geo, want_coef ? d_corr + slot * cs : nullptr, d_wacc, 1.0, kbar, 1, yi, yw,
                                  ystage, reduce_scratch(), stream);
      }
      if (d_hdot && st.clipped)
        launches += bk.re_dot(geo, d_hdot + slot, kbar, k[i], reduce_scratch(), stream);
    }
    const amp_t* ins[7];
    double w[7];
    ins[0] = lam; w[0] = 1.0;
    for (int i = 0; i < 6; ++i) { ins[i + 1] = yb[i]; w[i + 1] = 1.0; }
    launches += bk.lincomb(geo, lam, 7, ins, w, stream);
  }

  // Turn one slot's site correlations C_q[p][p'] into sample / time gradients.
  // dL = Re sum_{p,p'} dT_q[p][p'] * C_q[p][p'];  returns dL/dt_stage.
  double distribute(double ts, const cplx* C, double* g_det, double* g_amp) const {
    int nq = prog.nq, ns = prog.n_samples;
    double gd[kMaxQubits], ga[kMaxQubits], gb[kMaxQubits];
    for (int q = 0; q < nq; ++q) {
      if (prog.kind == PD_KET) {
        const cplx* c = C + q * 4;
        gd[q] = c[0].im;
        ga[q] = c[2].im + c[1].im;
        gb[q] = c[2].re - c[1].re;
      } else {
        const cplx* c = C + q * 16;
        gd[q] = ga[q] = gb[q] = 0.0;
        for (int a = 0; a < 2; ++a)
          for (int b = 0; b < 2; ++b) {
            int p = a * 2 + b, prow = (1 - a) * 2 + b, pcol = a * 2 + (1 - b);
            gd[q] += ((a == 0 ? 1.0 : 0.0) - (b == 0 ? 1.0 : 0.0)) * c[p * 4 + p].im;
            ga[q] += c[p * 4 + prow].im - c[p * 4 + pcol].im;
            gb[q] += (a == 1 ? 1.0 : -1.0) * c[p * 4 + prow].re + (b == 1 ? 1.0 : -1.0) * c[p * 4 + pcol].re;
          }
      }
    }
    if (ns < 2) return 0.0;
    Interp ix = interp_index(ts, prog.dt, ns);
    double x = (ts - ix.i1 * prog.dt) / prog.dt;
    double tbar = 0.0;
    for (int kdx = 0; kdx < prog.n_det(); ++kdx) {
      double g = 0.0;
      for (int q = 0; q < nq; ++q)
        if (prog.det_masks[kdx] >> q & 1) g += 2.0 * gd[q];
      const double* v = &prog.det_values[(size_t)kdx * ns];
      if (g_det) {
        g_det[(size_t)kdx * ns + ix.i1] += g * (1.0 - x);
        g_det[(size_t)kdx * ns + ix.i2] += g * x;
      }
      tbar += g * (v[ix.i2] - v[ix.i1]) / prog.dt;
    }
    for (int kdx = 0; kdx < prog.n_amp(); ++kdx) {
      double gre = 0.0, gim = 0.0;
      for (int q = 0; q < nq; ++q)
        if (prog.amp_masks[kdx] >> q & 1) { gre += ga[q]; gim += gb[q]; }
      const double* v = &prog.amp_values[(size_t)kdx * ns * 2];
      if (g_amp) {
        double* o1 = &g_amp[((size_t)kdx * ns + ix.i1) * 2];
        double* o2 = &g_amp[((size_t)kdx * ns + ix.i2) * 2];
        o1[0] += gre * (1.0 - x); o1[1] += gim * (1.0 - x);
        o2[0] += gre * x;         o2[1] += gim * x;
      }
      tbar += gre * (v[2 * ix.i2] - v[2 * ix.i1]) / prog.dt +
              gim * (v[2 * ix.i2 + 1] - v[2 * ix.i1 + 1]) / prog.dt;
    }
    return tbar;
  }

  // ---- Krylov --------------------------------------------------------------------------------
  // psi(t_k) = exp(-i (t_k - t_{k-1}) H(t_k)) psi(t_{k-1}), H frozen at the interval END
  // (KAT-confirmed K-B..K-E).  One Lanczos run per batch column.
  struct Lanczos {
    std::vector<double> alpha, beta;
    std::vector<cplx> w;
    double nrm = 0.0;
    int m = 0;
  };
  // Builds the Krylov basis of (H(t_eval), v0) in `basis` (m vectors of length dim), returns the
  // small-matrix data; out (nullable) = exp(-i*delta*H) v0 (sign=+1) or exp(+i*delta*H) v0 (-1).
  Lanczos lanczos_exp(const amp_t* v0, vec out, vec basis, int max_m, double t_eval, double delta,
                      double sign, const pd_options& o, void* stream) {
    size_t n = geo.dim;
    Lanczos lz;
    Geometry g1 = geo;
    g1.batch = 1;
    if ((bk.path == 0 || bk.path == 3) && bk.small_supported(g1, prog)) {
      // small registers: the recurrence runs on the device in chunks of iterations (small_ket.cuh,
      // k_small_lanczos); the stopping rule below is the same as in the launch-per-operation loop,
      // applied to the returned (alpha, beta)
      std::vector<double> al(max_m), be(max_m);
      const int chunk = 12;
      int done = 0;
      bool stop = false;
      while (!stop) {
        int j1 = std::min(max_m, done + chunk);
        launches += bk.small_lanczos(g1, prog, v0, basis, max_m, t_eval, done, j1, done > 0 ? be[done - 1] : 0.0,
                                     al.data(), be.data(), &lz.nrm, stream);
        if (done == 0 && lz.nrm == 0.0) {
          if (out) bk.zero(out, sizeof(amp_t) * n, stream);
          return lz;
        }
        for (int j = done; j < j1 && !stop; ++j) {
          lz.alpha.push_back(al[j]);
          double beta = be[j];
          lz.w = tridiag_expm_e1(lz.alpha, lz.beta, sign * delta);
          lz.m = j + 1;
          if (beta < o.norm_tolerance) { stop = true; break; }
          if (j >= 1) {
            const cplx& a = lz.w[j];
            const cplx& b = lz.w[j - 1];
            double est = (std::hypot(a.re, a.im) + std::hypot(b.re, b.im)) * beta * std::fabs(delta);
            if (est < o.exp_tolerance) { stop = true; break; }
          }
          if (j + 1 == max_m) { stop = true; break; }
          lz.beta.push_back(beta);
        }
        done = j1;
      }
      if (out) combine_basis(out, basis, lz.w, lz.m, lz.nrm, stream);
      return lz;
    }
    double* d_s = (double*)buf("kry_scal", sizeof(double) * 4);
    double hs[4];
    launches += bk.re_dot(g1, d_s, v0, v0, reduce_scratch(), stream);
    bk.d2h(hs, d_s, sizeof(double), stream);
    bk.sync(stream);
    lz.nrm = std::sqrt(hs[0]);
    if (lz.nrm == 0.0) {
      if (out) bk.zero(out, sizeof(amp_t) * n, stream);
      return lz;
    }
    {
      const amp_t* ins[1] = {v0};
      double w[1] = {1.0 / lz.nrm};
      launches += bk.lincomb(g1, basis, 1, ins, w, stream);
    }
    vec r = vbuf("kry_r");
    SiteOps so;
    prog.site_ops_ket(t_eval, 2, so);
    for (int j = 0; j < max_m; ++j) {
      vec vj = basis + (size_t)j * n;
      const amp_t* ins1[1] = {vj};
      double w1[1] = {1.0};
      launches += bk.stage_ket(g1, r, nullptr, 1, ins1, w1, so, scratch(stream), stream);
      launches += bk.re_dot(g1, d_s, vj, r, reduce_scratch(), stream);
      bk.d2h(hs, d_s, sizeof(double), stream);
      bk.sync(stream);
      lz.alpha.push_back(hs[0]);
      {
        const amp_t* ins[3] = {r, vj, j > 0 ? basis + (size_t)(j - 1) * n : vj};
        double w[3] = {1.0, -lz.alpha.back(), j > 0 ? -lz.beta.back() : 0.0};
        launches += bk.lincomb(g1, r, j > 0 ? 3 : 2, ins, w, stream);
      }
      launches += bk.re_dot(g1, d_s, r, r, reduce_scratch(), stream);
      bk.d2h(hs, d_s, sizeof(double), stream);
      bk.sync(stream);
      double beta = std::sqrt(hs[0]);
      lz.w = tridiag_expm_e1(lz.alpha, lz.beta, sign * delta);
      lz.m = j + 1;
      if (beta < o.norm_tolerance) break;
      if (j >= 1) {
        const cplx& a = lz.w[j];
        const cplx& b = lz.w[j - 1];
        double est = (std::hypot(a.re, a.im) + std::hypot(b.re, b.im)) * beta * std::fabs(delta);
        if (est < o.exp_tolerance) break;
      }
      if (j + 1 == max_m) break;
      lz.beta.push_back(beta);
      const amp_t* ins[1] = {r};
      double w[1] = {1.0 / beta};
      launches += bk.lincomb(g1, basis + (size_t)(j + 1) * n, 1, ins, w, stream);
    }
    if (out) combine_basis(out, basis, lz.w, lz.m, lz.nrm, stream);
    return lz;
  }
  // out = scale * sum_j w_j basis_j  (complex weights)
  void combine_basis(vec out, const amp_t* basis, const std::vector<cplx>& w, int m, double scale,
                     void* stream) {
    Geometry g1 = geo;
    g1.batch = 1;
    std::vector<cplx> ws(m);
    for (int j = 0; j < m; ++j) ws[j] = scale * w[j];
    launches += bk.lincomb_c(g1, out, m, basis, geo.dim, ws.data(), stream);
  }

  void forward_krylov(const pd_options& o, const amp_t* state0, const double* tsave, int n_t,
                      amp_t* states, Tape* tape, void* stream) {
    if (prog.kind != PD_KET) throw Error(PD_ERR_INVALID, "KRYLOV_SE needs a ket plan");
    size_t n = geo.dim;
    int max_m = std::max(2, (int)o.max_krylov);
    vec basis = (vec)buf("kry_basis", sizeof(amp_t) * n * (size_t)max_m);
    bk.d2d(states, state0, sizeof(amp_t) * L, stream);
    for (int kk = 1; kk < n_t; ++kk) {
      double delta = tsave[kk] - tsave[kk - 1];
      amp_t* dst = states + (size_t)kk * L;
      const amp_t* src = states + (size_t)(kk - 1) * L;
      if (!(delta > 0.0)) { bk.d2d(dst, src, sizeof(amp_t) * L, stream); continue; }
      for (int b = 0; b < geo.batch; ++b)
        lanczos_exp(src + (size_t)b * n, dst + (size_t)b * n, basis, max_m, tsave[kk], delta, 1.0, o, stream);
      if (tape) tape->ksegs.push_back({kk, delta, tsave[kk]});
    }
    bk.sync(stream);
  }

  // Adjoint of the exact exponential per interval:
  //   lam_{k-1} = exp(+i D H) lam_k,
  //   dL/dtheta = Re <lam_k, d/dtheta exp(-i D H) psi_{k-1}>
  //             = D * int_0^1 Re < lam(s) , -i (dH/dtheta) psi(s) > ds,
  //   psi(s) = exp(-i s D H) psi_{k-1},  lam(s) = exp(+i (1-s) D H) lam_k,
  // with both curves evaluated inside their Krylov subspaces and the integral by
  // Gauss-Legendre quadrature.  Matches tape autograd through Lanczos to the Krylov tolerance.
  void backward_krylov(Tape& tape, const amp_t* states, const amp_t* gstates, double* g_det,
                       double* g_amp, double* g_pair, double* g_tsave, amp_t* g_state0,
                       bool want_coef, void* stream) {
    static const double gx[12] = {-0.9815606342467192, -0.9041172563704749, -0.7699026741943047,
                                  -0.5873179542866175, -0.3678314989981802, -0.1252334085114689,
                                  0.1252334085114689,  0.3678314989981802,  0.5873179542866175,
                                  0.7699026741943047,  0.9041172563704749,  0.9815606342467192};
    static const double gw[12] = {0.0471753363865118, 0.1069393259953184, 0.1600783285433462,
                                  0.2031674267230659, 0.2334925365383548, 0.2491470458134028,
                                  0.2491470458134028, 0.2334925365383548, 0.2031674267230659,
                                  0.1600783285433462, 0.1069393259953184, 0.0471753363865118};
    const int nquad = 12;
    int n_t = (int)tape.tsave.size();
    size_t n = geo.dim;
    const pd_options& o = tape.opt;
    int max_m = std::max(2, (int)o.max_krylov);
    Geometry g1 = geo;
    g1.batch = 1;
    vec lam = vbuf("lam");
    if (gstates) bk.d2d(lam, gstates + (size_t)(n_t - 1) * L, sizeof(amp_t) * L, stream);
    else bk.zero(lam, sizeof(amp_t) * L, stream);
    vec basisV = (vec)buf("kry_basis", sizeof(amp_t) * n * (size_t)max_m);
    vec basisW = (vec)buf("kry_basis2", sizeof(amp_t) * n * (size_t)max_m);
    vec ps = vbuf("kry_ps"), ls = vbuf("kry_ls"), lnew = vbuf("kry_lnew");
    int cs = corr_stride();
    double* d_wacc = nullptr;
    if (g_pair) {
      d_wacc = (double*)buf("wacc", sizeof(double) * n);
      bk.zero(d_wacc, sizeof(double) * n, stream);
    }
    size_t n_seg = tape.ksegs.size();
    size_t n_slots = n_seg * geo.batch * nquad;
    cplx* d_corr = (cplx*)buf("corr", sizeof(cplx) * std::max<size_t>(1, n_slots * cs));
    struct KSlot { double t_eval; double weight; int interval; };
    std::vector<KSlot> kslots(n_slots, KSlot{0.0, 0.0, 0});
    double* d_h = (double*)buf("kry_scal", sizeof(double) * 4);
    std::vector<int> seg_of(n_t, -1);
    for (size_t si = 0; si < n_seg; ++si) seg_of[tape.ksegs[si].interval] = (int)si;
    size_t slot = 0;
    for (int kk = n_t - 1; kk >= 1; --kk) {
      if (seg_of[kk] >= 0) {
        const auto& sg = tape.ksegs[seg_of[kk]];
        SiteOps so_rhs;
        prog.site_ops_ket(sg.t_eval, 0, so_rhs);
        for (int b = 0; b < geo.batch; ++b) {
          const amp_t* psi_prev = states + (size_t)(kk - 1) * L + (size_t)b * n;
          const amp_t* psi_k = states + (size_t)kk * L + (size_t)b * n;
          vec lam_b = lam + (size_t)b * n;
          if (g_tsave) {
            // through delta = t_k - t_{k-1}:  Re<lam_k, -i H psi_k>
            const amp_t* ins1[1] = {psi_k};
            double w1[1] = {1.0};
            launches += bk.stage_ket(g1, ps, nullptr, 1, ins1, w1, so_rhs, scratch(stream), stream);
            launches += bk.re_dot(g1, d_h, lam_b, ps, reduce_scratch(), stream);
            double hv;
            bk.d2h(&hv, d_h, sizeof(double), stream);
            bk.sync(stream);
            g_tsave[kk] += hv;
            g_tsave[kk - 1] -= hv;
          }
          Lanczos lv = lanczos_exp(psi_prev, nullptr, basisV, max_m, sg.t_eval, sg.delta, 1.0, o, stream);
          Lanczos lw = lanczos_exp(lam_b, lnew, basisW, max_m, sg.t_eval, sg.delta, -1.0, o, stream);
          for (int qd = 0; qd < nquad; ++qd, ++slot) {
            if (!(want_coef || d_wacc) || lv.m == 0 || lw.m == 0) continue;
            double s = 0.5 * (gx[qd] + 1.0);
            double wq = 0.5 * gw[qd] * sg.delta;
            std::vector<cplx> wv = tridiag_expm_e1(lv.alpha, lv.beta, s * sg.delta);
            std::vector<cplx> ww = tridiag_expm_e1(lw.alpha, lw.beta, -(1.0 - s) * sg.delta);
            combine_basis(ps, basisV, wv, lv.m, lv.nrm, stream);
            combine_basis(ls, basisW, ww, lw.m, lw.nrm, stream);
            launches += bk.corr(g1, d_corr + slot * cs, d_wacc, wq, ls, ps, reduce_scratch(), stream);
            kslots[slot] = {sg.t_eval, wq, kk};
          }
          bk.d2d(lam_b, lnew, sizeof(amp_t) * n, stream);
        }
      }
      if (gstates) {
        const amp_t* ins[2] = {lam, gstates + (size_t)(kk - 1) * L};
        double w[2] = {1.0, 1.0};
        launches += bk.lincomb(geo, lam, 2, ins, w, stream);
      }
    }
    if (g_state0) bk.d2d(g_state0, lam, sizeof(amp_t) * L, stream);
    if (want_coef && n_slots > 0) {
      std::vector<cplx> h_corr(n_slots * cs);
      bk.d2h(h_corr.data(), d_corr, sizeof(cplx) * n_slots * cs, stream);
      bk.sync(stream);
      std::vector<cplx> scaled(cs);
      for (size_t s2 = 0; s2 < n_slots; ++s2) {
        const KSlot& ks = kslots[s2];
        if (ks.weight == 0.0) continue;
        for (int i = 0; i < cs; ++i) scaled[i] = ks.weight * h_corr[s2 * cs + i];
        double tbar = distribute(ks.t_eval, scaled.data(), g_det, g_amp);
        if (g_tsave) g_tsave[ks.interval] += tbar;  // H is evaluated at the interval end t_k
      }
    }
    if (g_pair) {
      double* d_pair = (double*)buf("pair_out", sizeof(double) * (size_t)prog.nq * prog.nq);
      launches += bk.pair_reduce(geo, d_pair, d_wacc, stream);
      bk.d2h(g_pair, d_pair, sizeof(double) * (size_t)prog.nq * prog.nq, stream);
    }
    bk.sync(stream);
  }
};

}  // namespace pd
