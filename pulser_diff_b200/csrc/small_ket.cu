// Host side of the small-register ket family (kernels: small_ket.cuh, instantiated in
// small_ket_fwd.cu / small_ket_bwd1.cu / small_ket_bwd2.cu so that they compile in parallel).
#include <chrono>

#include "small_ket.cuh"

namespace pd {
namespace sk {
void launch_forward(int nq, const SkFwd& P, int nC, cudaStream_t st);
void launch_backward(int nq, const SkBwd& P, int nC, cudaStream_t st);
void launch_lanczos(int nq, const SkLanczos& P, int nC, cudaStream_t st);
}  // namespace sk
using namespace sk;

// ---- host side ------------------------------------------------------------------------------
struct SmallKetState {
  // device copies of the pulse program
  unsigned long long *d_dm = nullptr, *d_am = nullptr;
  double *d_dv = nullptr, *d_av = nullptr;
  size_t cap_dm = 0, cap_am = 0, cap_dv = 0, cap_av = 0;
  uint64_t uploaded_version = 0;   // Program::version of the single-unit tables currently on the device
  // workspace
  void* ws[20] = {};
  size_t ws_cap[20] = {};
  // stage tape of the most recent recorded forward sweep (slots 14/15) and its identity
  uint64_t tape_gen = 0;
  std::vector<int> tape_steps;   // accepted steps per unit of that sweep
  std::vector<int> tape_attempts;
  int tape_cap = 0;
  bool steps_on_device = false;  // batches: step lists live in ws[16] / ws[17] (k_compact_log)
  ~SmallKetState() {
    cudaFree(d_dm); cudaFree(d_am); cudaFree(d_dv); cudaFree(d_av);
    for (auto p : ws) cudaFree(p);
  }
  void* get(int slot, size_t bytes) {
    if (ws_cap[slot] < bytes) {
      cudaFree(ws[slot]);
      ws[slot] = nullptr;
      PD_CUDA_CHECK(cudaMalloc(&ws[slot], std::max<size_t>(bytes, 256)));
      ws_cap[slot] = bytes;
    }
    return ws[slot];
  }
};

SmallKetState* small_ket_create() { return new SmallKetState(); }
void small_ket_destroy(SmallKetState* s) { delete s; }

static int small_env(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

// number of CTAs (SK_T real lanes each); false if the register does not fit one cooperative launch
static bool small_shape(size_t L, int batch, int& nC) {
  (void)batch;
  // a unit that spans several CTAs needs them all co-resident (cooperative launch): two CTAs of
  // 128 threads fit an SM with this kernel's registers and shared memory
  static int per_device[64] = {};          // a process may drive several devices
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (per_device[dev] == 0) {
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) sms = 0;
    per_device[dev] = std::max(1, std::min(SK_MAXC, 2 * sms));
  }
  const int max_coresident = per_device[dev];
  size_t c = (2 * L + SK_T - 1) / SK_T;
  if (c > 1 && c > (size_t)max_coresident) return false;
  nC = (int)c;
  return true;
}

bool small_ket_supported(const Geometry& g, const Program& prog) {
  if (small_env("PD_SMALL_DISABLE", 0)) return false;
  if (g.kind != PD_KET || g.nq > SK_MAXQ || g.batch > SK_MAXB) return false;
  if (prog.n_det() > SK_MAXTERMS || prog.n_amp() > SK_MAXTERMS) return false;
  int nC;
  return small_shape(g.dim * (size_t)g.batch, g.batch, nC);
}

// Uploads masks and the coefficient tables of n_units units (dv: [U][n_det][ns], av: [U][n_amp][ns][2]).
static void upload_prog(SmallKetState& S, const Program& prog, const Geometry& g, int n_units, const double* dv,
                        const double* av, SkProg& o, cudaStream_t st) {
  auto up = [&](auto*& dptr, size_t& cap, const void* src, size_t bytes) {
    if (cap < bytes) {
      cudaFree(dptr);
      dptr = nullptr;
      PD_CUDA_CHECK(cudaMalloc((void**)&dptr, std::max<size_t>(bytes, 64)));
      cap = bytes;
    }
    if (bytes) PD_CUDA_CHECK(copy_in(dptr, src, bytes, st));   // tables may already be on the device
  };
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "mask width");
  const size_t ns = (size_t)prog.n_samples;
  // the plan's own tables stay on the device until the program changes (Program::version)
  const bool own = n_units == 1 && dv == prog.det_values.data() && av == prog.amp_values.data();
  if (!(own && S.uploaded_version == prog.version && prog.version != 0)) {
    up(S.d_dm, S.cap_dm, prog.det_masks.data(), prog.det_masks.size() * 8);
    up(S.d_am, S.cap_am, prog.amp_masks.data(), prog.amp_masks.size() * 8);
    up(S.d_dv, S.cap_dv, dv, (size_t)n_units * prog.n_det() * ns * 8);
    up(S.d_av, S.cap_av, av, (size_t)n_units * prog.n_amp() * ns * 16);
    S.uploaded_version = own ? prog.version : 0;
  }
  o.nq = prog.nq; o.n_samples = prog.n_samples; o.n_det = prog.n_det(); o.n_amp = prog.n_amp();
  o.dt = prog.dt;
  o.det_masks = S.d_dm; o.det_values = S.d_dv; o.amp_masks = S.d_am; o.amp_values = S.d_av;
  o.diag = g.diag;
}

bool small_ket_units_supported(const Geometry& g, const Program& prog) {
  int nC;
  return small_ket_supported(g, prog) && small_shape(g.dim * (size_t)g.batch, g.batch, nC) && nC == 1;
}

// Whole forward evolution of n_units independent parameter sets (same register, masks and time
// grid; per-unit coefficient tables dv/av on the host, null = the plan's own for a single unit).
// y0: [U][L] states at tsave[0]; states: [U][n_t][L].  k1 = f(t0, y0) and the initial step are
// computed on the device.  Appends every attempt of unit u to records[u].  Returns the launches.
int small_ket_forward(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                      const pd_options& o, int n_units, const cplx* y0, const double* dv, const double* av,
                      const double* tsave, int n_t, cplx* states,
                      std::vector<std::vector<pd_step_record>>& records, bool want_tape,
                      uint64_t* tape_gen_out, cudaStream_t st) {
  const size_t L = g.dim * (size_t)g.batch;
  const size_t U = (size_t)n_units;
  int nC;
  if (!small_shape(L, g.batch, nC)) throw Error(PD_ERR_STATE, "small_ket_forward: unsupported shape");
  // PD_TIMING=1: host-side phase times of this call on stderr
  static const bool timing = getenv("PD_TIMING") != nullptr;
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t_in = timing ? now() : 0.0;
  double t_tape = 0.0, t_launch = 0.0, t_done = 0.0;
  if (n_units > 1 && nC != 1)
    throw Error(PD_ERR_INVALID, "batches of parameter sets need 2 * batch * 2^N <= 128 (one CTA per set)");
  if (tape_gen_out) *tape_gen_out = 0;
  records.assign(U, {});
  SkFwd P{};
  upload_prog(S, prog, g, n_units, dv ? dv : prog.det_values.data(), av ? av : prog.amp_values.data(), P.prog, st);
  fill_tab(tab, P.tab);
  P.batch = g.batch; P.nC = nC; P.dim = g.dim; P.L = L;
  P.n_units = n_units;
  P.det_stride = (size_t)prog.n_det() * prog.n_samples;
  P.amp_stride = (size_t)prog.n_amp() * prog.n_samples * 2;
  P.atol = o.atol; P.rtol = o.rtol; P.safety = o.safety_factor; P.minf = o.min_factor; P.maxf = o.max_factor;
  P.max_steps = o.max_steps;
  P.n_replay = o.n_replay;
  const int log_cap = n_units == 1 ? (1 << 15) : (int)std::max<size_t>(256, std::min<size_t>(1 << 15, ((size_t)64 << 20) / 40 / U));
  double* d_ts = (double*)S.get(0, sizeof(double) * n_t);
  PD_CUDA_CHECK(copy_h2d(d_ts, tsave, sizeof(double) * n_t, st));
  P.tsave = d_ts; P.n_t = n_t;
  P.y_io = (cplx*)S.get(1, sizeof(cplx) * L * U);
  P.k0_io = (cplx*)S.get(2, sizeof(cplx) * L * U);
  PD_CUDA_CHECK(cudaMemcpyAsync(P.y_io, y0, sizeof(cplx) * L * U, cudaMemcpyDeviceToDevice, st));
  P.states = states;
  const size_t ys_bytes = sizeof(uint4) * 2 * (2 * L) * U, red_bytes = sizeof(uint4) * 2 * nC * g.batch * U;
  P.YS = (uint4*)S.get(3, ys_bytes);
  P.red = (uint4*)S.get(4, red_bytes);
  P.log = (pd_step_record*)S.get(5, sizeof(pd_step_record) * log_cap * U);
  P.log_cap = log_cap;
  P.resume = (SkResume*)S.get(6, sizeof(SkResume) * U);
  P.abort_flag = (int*)S.get(10, 64);
  if (o.n_replay > 0 && n_units > 1) throw Error(PD_ERR_INVALID, "replayed step sequences are per evolution: not available for batches");
  if (o.n_replay > 0) {
    double* d_rd = (double*)S.get(7, sizeof(double) * o.n_replay);
    unsigned char* d_rc = (unsigned char*)S.get(8, (size_t)o.n_replay);
    PD_CUDA_CHECK(copy_h2d(d_rd, o.replay_dt, sizeof(double) * o.n_replay, st));
    PD_CUDA_CHECK(copy_h2d(d_rc, o.replay_clipped, (size_t)o.n_replay, st));
    P.replay_dt = d_rd; P.replay_clipped = d_rc;
  }
  if (timing) t_tape = now();
  if (want_tape) {
    // per accepted step: 6 stage inputs + 6 slopes; sized from the time grid, bounded by memory
    const size_t per_step = 6 * L * sizeof(cplx) * U;
    const size_t want_steps = (size_t)16 * n_t + 256;
    size_t steps_cap = S.ws_cap[14] / per_step;            // what the plan already owns
    if (steps_cap < want_steps) {
      // grow (rare): only now ask the driver how much is free -- cudaMemGetInfo stalls for tens of
      // milliseconds every so often and must stay out of the steady state
      size_t fr = 0, tot = 0;
      if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) fr = (size_t)4 << 30;
      const size_t budget = std::max<size_t>((size_t)3 << 29, fr / 8) + S.ws_cap[14];
      steps_cap = std::max(steps_cap, std::min(want_steps, budget / per_step));
    }
    if (steps_cap >= 1) {
      P.tapeY = (double*)S.get(14, steps_cap * per_step);
      P.tapeK = (double*)S.get(15, steps_cap * per_step);
      P.tape_cap = (int)std::min<size_t>(steps_cap, (size_t)1 << 30);
    }
    S.tape_steps.clear();
    S.tape_cap = P.tape_cap;
  }
  // batches keep their step lists on the device (ws[16]: steps, ws[17]: accepted / attempted counts)
  const bool dev_steps = U > 1 && P.tapeY != nullptr;
  SkStep* d_steps = nullptr;
  int *d_nsteps = nullptr, *d_natt = nullptr;
  if (dev_steps) {
    d_steps = (SkStep*)S.get(16, sizeof(SkStep) * U * (size_t)P.tape_cap);
    d_nsteps = (int*)S.get(17, sizeof(int) * U * 2);
    d_natt = d_nsteps + U;
    PD_CUDA_CHECK(cudaMemsetAsync(d_nsteps, 0, sizeof(int) * U * 2, st));
  }
  std::vector<SkResume> rs(U);
  for (auto& r : rs) {
    r = SkResume{};
    r.t = tsave[0]; r.error = 1.0; r.cache_err = 1.0;
    r.kk = -1; r.n_acc = 0; r.tape_ok = P.tapeY ? 1 : 0;
  }
  PD_CUDA_CHECK(copy_h2d(P.resume, rs.data(), sizeof(SkResume) * U, st));
  int launches = 0;
  std::vector<pd_step_record> chunk;
  for (;;) {
    // the exchange lines carry phase numbers that restart at 1 in every launch
    PD_CUDA_CHECK(cudaMemsetAsync(P.YS, 0, ys_bytes, st));
    PD_CUDA_CHECK(cudaMemsetAsync(P.red, 0, red_bytes, st));
    PD_CUDA_CHECK(cudaMemsetAsync(P.abort_flag, 0, 64, st));
    if (timing) t_launch = now();
    sk::launch_forward(prog.nq, P, nC, st);
    ++launches;
    PD_CUDA_CHECK(copy_d2h(rs.data(), P.resume, sizeof(SkResume) * U, st));
    PD_CUDA_CHECK(cudaStreamSynchronize(st));
    if (timing) t_done = now();
    size_t max_rec = 0;
    for (auto& r : rs) max_rec = std::max<size_t>(max_rec, (size_t)r.n_rec);
    if (max_rec > 0 && dev_steps) {
      k_compact_log<<<(unsigned)((U + 127) / 128), 128, 0, st>>>(P.log, log_cap, P.resume, d_steps, P.tape_cap,
                                                                 d_nsteps, d_natt, (int)U);
      PD_CUDA_CHECK(cudaGetLastError());
      ++launches;
    } else if (max_rec > 0) {
      if (U == 1) {
        chunk.resize(rs[0].n_rec);
        PD_CUDA_CHECK(copy_d2h(chunk.data(), P.log, sizeof(pd_step_record) * rs[0].n_rec, st));
        PD_CUDA_CHECK(cudaStreamSynchronize(st));
        records[0].insert(records[0].end(), chunk.begin(), chunk.end());
      } else {
        chunk.resize((size_t)log_cap * U);
        PD_CUDA_CHECK(copy_d2h(chunk.data(), P.log, sizeof(pd_step_record) * log_cap * U, st));
        PD_CUDA_CHECK(cudaStreamSynchronize(st));
        for (size_t u = 0; u < U; ++u)
          records[u].insert(records[u].end(), chunk.begin() + u * log_cap, chunk.begin() + u * log_cap + rs[u].n_rec);
      }
    }
    bool again = false;
    for (auto& r : rs) {
      if (r.status == 1) again = true;
      if (r.status == 2) throw Error(PD_ERR_MAX_STEPS, "max_steps reached");
      if (r.status == 3) throw Error(PD_ERR_STATE, "non-finite error norm in DP5 step");
      if (r.status == 4) throw Error(PD_ERR_INVALID, "replay sequence too short");
      if (r.status == 5) throw Error(PD_ERR_STATE, "small_ket_forward: exchange poll timed out");
    }
    if (!again) break;
  }
  if (timing)
    fprintf(stderr, "[pd] small forward: setup %.3f ms, tape sizing+rest %.3f ms, last launch->sync %.3f ms, logs %.3f ms\n",
            t_tape - t_in, t_launch - t_tape, t_done - t_launch, now() - t_done);
  if (P.tapeY) {
    bool ok = true;
    for (auto& r : rs) ok = ok && r.tape_ok;
    if (ok) {
      S.tape_steps.resize(U);
      for (size_t u = 0; u < U; ++u) S.tape_steps[u] = rs[u].n_acc;
      S.steps_on_device = dev_steps;
      if (dev_steps) {
        S.tape_attempts.resize(U);
        PD_CUDA_CHECK(copy_d2h(S.tape_attempts.data(), d_natt, sizeof(int) * U, st));
        PD_CUDA_CHECK(cudaStreamSynchronize(st));
      }
      ++S.tape_gen;
      if (tape_gen_out) *tape_gen_out = S.tape_gen;
    }
  }
  return launches;
}

size_t small_ket_nred(const Program& prog) { return (size_t)prog.n_det() + 2 * (size_t)prog.n_amp() + 1; }

// Whole adjoint sweep of n_units units over the stage tape recorded by the forward sweep `tape_gen`.
// steps[u] = accepted steps of unit u.  slot_sums[u] (host, [n_steps_u*6][nred]) receives per slot
// the per-term sums (det terms: sum_q gd_q; amp terms: sum_q ga_q, sum_q gb_q; last: Re<kbar, k_i>).
// gstates: [U][n_t][L] or null; lam_out: [U][L]; d_wacc only for a single unit.
// Returns the number of launches, or 0 if that tape is gone (another forward sweep ran on the plan
// since) or was incomplete: the caller then falls back to the stage-by-stage adjoint.
int small_ket_backward(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                       const std::vector<double>& tsave, int n_units, const double* dv, const double* av,
                       const std::vector<std::vector<SkStepHost>>& steps, uint64_t tape_gen,
                       const cplx* gstates, bool want_coef, double* d_wacc, cplx* lam_out,
                       std::vector<std::vector<double>>& slot_sums, cudaStream_t st) {
  const size_t L = g.dim * (size_t)g.batch;
  const size_t U = (size_t)n_units;
  int nC;
  if (!small_shape(L, g.batch, nC)) throw Error(PD_ERR_STATE, "small_ket_backward: unsupported shape");
  const int n_t = (int)tsave.size();
  const size_t nred = small_ket_nred(prog);
  if (tape_gen == 0 || tape_gen != S.tape_gen || S.tape_steps.size() != U) return 0;
  size_t max_steps = 1;
  for (size_t u = 0; u < U; ++u) {
    if (S.tape_steps[u] != (int)steps[u].size()) return 0;
    max_steps = std::max(max_steps, steps[u].size());
  }
  if (U * max_steps * 6 * nC * nred * sizeof(double) > ((size_t)2 << 30)) return 0;
  SkBwd P{};
  upload_prog(S, prog, g, n_units, dv ? dv : prog.det_values.data(), av ? av : prog.amp_values.data(), P.prog, st);
  fill_tab(tab, P.tab);
  P.batch = g.batch; P.nC = nC; P.dim = g.dim; P.L = L; P.n_t = n_t;
  P.n_units = n_units; P.max_steps = (int)max_steps; P.tape_cap = S.tape_cap;
  P.det_stride = (size_t)prog.n_det() * prog.n_samples;
  P.amp_stride = (size_t)prog.n_amp() * prog.n_samples * 2;
  std::vector<SkStep> hs(U * max_steps);
  std::vector<int> hn(U);
  for (size_t u = 0; u < U; ++u) {
    hn[u] = (int)steps[u].size();
    for (size_t i = 0; i < steps[u].size(); ++i)
      hs[u * max_steps + i] = {steps[u][i].t, steps[u][i].dt, steps[u][i].interval, steps[u][i].clipped};
  }
  SkStep* d_steps = (SkStep*)S.get(9, sizeof(SkStep) * hs.size());
  int* d_n = (int*)S.get(7, sizeof(int) * U);
  PD_CUDA_CHECK(copy_h2d(d_steps, hs.data(), sizeof(SkStep) * hs.size(), st));
  PD_CUDA_CHECK(copy_h2d(d_n, hn.data(), sizeof(int) * U, st));
  P.steps = d_steps; P.unit_steps = d_n; P.n_steps = (int)steps[0].size();
  P.gstates = gstates;
  P.tapeY = (const double*)S.ws[14];
  P.tapeK = (const double*)S.ws[15];
  const size_t kb_bytes = sizeof(uint4) * 2 * (2 * L) * U;
  P.KB = (uint4*)S.get(3, kb_bytes);
  PD_CUDA_CHECK(cudaMemsetAsync(P.KB, 0, kb_bytes, st));
  const size_t n_part = U * max_steps * 6 * nC * nred;
  P.slotpart = (double*)S.get(12, sizeof(double) * n_part);
  double* d_we = d_wacc ? (double*)S.get(13, sizeof(double) * L * U) : nullptr;
  if (d_wacc && U != 1) throw Error(PD_ERR_INVALID, "pair-coupling gradients are not available for batches of parameter sets");
  P.wacc_elem = d_we;
  P.lam_out = lam_out;
  P.want_coef = want_coef ? 1 : 0;
  P.abort_flag = (int*)S.get(10, 64);
  PD_CUDA_CHECK(cudaMemsetAsync(P.abort_flag, 0, 64, st));
  sk::launch_backward(prog.nq, P, nC, st);
  int launches = 1;
  if (d_wacc) {
    k_fold_columns<<<(unsigned)((g.dim + 255) / 256), 256, 0, st>>>(d_we, d_wacc, g.dim, g.batch);
    ++launches;
  }
  slot_sums.assign(U, {});
  if (want_coef) {
    const size_t n_out = U * max_steps * 6 * nred;
    const double* d_src = P.slotpart;
    if (nC > 1) {
      double* d_out = (double*)S.get(11, sizeof(double) * n_out);
      k_sum_slots<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(P.slotpart, d_out, U * max_steps * 6, nC, (int)nred);
      ++launches;
      d_src = d_out;
    }
    std::vector<double> all(n_out);
    PD_CUDA_CHECK(copy_d2h(all.data(), d_src, sizeof(double) * n_out, st));
    PD_CUDA_CHECK(cudaStreamSynchronize(st));
    for (size_t u = 0; u < U; ++u)
      slot_sums[u].assign(all.begin() + u * max_steps * 6 * nred,
                          all.begin() + u * max_steps * 6 * nred + steps[u].size() * 6 * nred);
  }
  PD_CUDA_CHECK(cudaStreamSynchronize(st));
  PD_CUDA_CHECK(cudaGetLastError());
  {
    int ab = 0;
    PD_CUDA_CHECK(copy_d2h(&ab, P.abort_flag, sizeof(int), st));
    PD_CUDA_CHECK(cudaStreamSynchronize(st));
    if (ab) throw Error(PD_ERR_STATE, "small_ket_backward: exchange poll timed out");
  }
  return launches;
}


// Lanczos iterations [j0, j1) of one column vector on the device (KRYLOV_SE); fills basis[j0+1..j1]
// (capped at max_m - 1), and alpha[j0..j1) / beta[j0..j1) on the host.  j0 = 0 normalises v0 into
// basis[0] and returns |v0| in *nrm.  One launch.
int small_ket_lanczos(SmallKetState& S, const Geometry& g1, const Program& prog, const cplx* v0, cplx* basis,
                      int max_m, double t_eval, int j0, int j1, double beta_prev, double* alpha_host,
                      double* beta_host, double* nrm, cudaStream_t st) {
  int nC;
  if (!small_shape(g1.dim, 1, nC)) throw Error(PD_ERR_STATE, "small_ket_lanczos: unsupported shape");
  SkLanczos P{};
  upload_prog(S, prog, g1, 1, prog.det_values.data(), prog.amp_values.data(), P.prog, st);
  P.nC = nC; P.dim = g1.dim; P.t_eval = t_eval; P.j0 = j0; P.j1 = j1; P.max_m = max_m; P.beta_prev = beta_prev;
  P.basis = basis; P.v0 = v0;
  double* d_ab = (double*)S.get(18, sizeof(double) * (2 * (size_t)max_m + 1));
  P.alpha = d_ab; P.beta = d_ab + max_m; P.nrm_out = d_ab + 2 * max_m;
  const size_t ys_bytes = sizeof(uint4) * 2 * (2 * g1.dim), red_bytes = sizeof(uint4) * 2 * nC;
  P.YS = (uint4*)S.get(3, ys_bytes);
  P.red = (uint4*)S.get(4, red_bytes);
  P.abort_flag = (int*)S.get(10, 64);
  PD_CUDA_CHECK(cudaMemsetAsync(P.YS, 0, ys_bytes, st));
  PD_CUDA_CHECK(cudaMemsetAsync(P.red, 0, red_bytes, st));
  PD_CUDA_CHECK(cudaMemsetAsync(P.abort_flag, 0, 64, st));
  sk::launch_lanczos(prog.nq, P, nC, st);
  PD_CUDA_CHECK(copy_d2h(alpha_host + j0, P.alpha + j0, sizeof(double) * (j1 - j0), st));
  PD_CUDA_CHECK(copy_d2h(beta_host + j0, P.beta + j0, sizeof(double) * (j1 - j0), st));
  if (j0 == 0) PD_CUDA_CHECK(copy_d2h(nrm, P.nrm_out, sizeof(double), st));
  int ab = 0;
  PD_CUDA_CHECK(copy_d2h(&ab, P.abort_flag, sizeof(int), st));
  PD_CUDA_CHECK(cudaStreamSynchronize(st));
  if (ab) throw Error(PD_ERR_STATE, "small_ket_lanczos: exchange poll timed out");
  return 1;
}

// accepted / attempted steps of unit u of the batch recorded by forward sweep `tape_gen` (-1: unknown)
void small_ket_unit_counts(SmallKetState& S, uint64_t tape_gen, int unit, int* accepted, int* attempts) {
  *accepted = *attempts = -1;
  if (tape_gen == 0 || tape_gen != S.tape_gen || unit < 0 || (size_t)unit >= S.tape_steps.size()) return;
  *accepted = S.tape_steps[unit];
  if ((size_t)unit < S.tape_attempts.size()) *attempts = S.tape_attempts[unit];
}

// Adjoint sweep of a batch whose step lists stayed on the device: kernel, then the gradient scatter
// on the device; only the sample gradients come back.  g_det: [U][n_det][ns], g_amp: [U][n_amp][ns][2]
// (host, nullable).  Returns launches, 0 if that tape is gone.
int small_ket_backward_units(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                             const std::vector<double>& tsave, int n_units, const double* dv, const double* av,
                             uint64_t tape_gen, const cplx* gstates, cplx* lam_out, double* g_det, double* g_amp,
                             cudaStream_t st) {
  const size_t L = g.dim * (size_t)g.batch;
  const size_t U = (size_t)n_units;
  int nC;
  if (!small_shape(L, g.batch, nC) || nC != 1) throw Error(PD_ERR_STATE, "small_ket_backward_units: unsupported shape");
  if (tape_gen == 0 || tape_gen != S.tape_gen || S.tape_steps.size() != U || !S.steps_on_device) return 0;
  const int n_t = (int)tsave.size();
  const size_t nred = small_ket_nred(prog);
  const int ns = prog.n_samples, n_det = prog.n_det(), n_amp = prog.n_amp();
  size_t max_steps = 1;
  for (size_t u = 0; u < U; ++u) max_steps = std::max<size_t>(max_steps, (size_t)S.tape_steps[u]);
  const size_t acc_bytes = sizeof(double) * (size_t)(n_det + 2 * n_amp) * ns;
  if (acc_bytes > 40 * 1024) return 0;
  SkBwd P{};
  upload_prog(S, prog, g, n_units, dv ? dv : prog.det_values.data(), av ? av : prog.amp_values.data(), P.prog, st);
  fill_tab(tab, P.tab);
  P.batch = g.batch; P.nC = nC; P.dim = g.dim; P.L = L; P.n_t = n_t;
  P.n_units = n_units; P.tape_cap = S.tape_cap;
  P.max_steps = S.tape_cap;                 // stride of the device step lists
  P.det_stride = (size_t)n_det * ns;
  P.amp_stride = (size_t)n_amp * ns * 2;
  P.steps = (const SkStep*)S.ws[16];
  P.unit_steps = (const int*)S.ws[17];
  P.n_steps = 0;
  P.gstates = gstates;
  P.tapeY = (const double*)S.ws[14];
  P.tapeK = (const double*)S.ws[15];
  const size_t kb_bytes = sizeof(uint4) * 2 * (2 * L) * U;
  P.KB = (uint4*)S.get(3, kb_bytes);
  PD_CUDA_CHECK(cudaMemsetAsync(P.KB, 0, kb_bytes, st));
  // slot sums: stride tape_cap steps per unit (the kernel indexes with max_steps = tape_cap)
  const size_t n_part = U * (size_t)S.tape_cap * 6 * nC * nred;
  if (n_part * sizeof(double) > ((size_t)8 << 30)) return 0;
  P.slotpart = (double*)S.get(12, sizeof(double) * n_part);
  P.wacc_elem = nullptr;
  P.lam_out = lam_out;
  const bool want_coef = g_det || g_amp;
  P.want_coef = want_coef ? 1 : 0;
  P.abort_flag = (int*)S.get(10, 64);
  PD_CUDA_CHECK(cudaMemsetAsync(P.abort_flag, 0, 64, st));
  sk::launch_backward(prog.nq, P, nC, st);
  int launches = 1;
  if (want_coef) {
    double* d_gd = (double*)S.get(11, sizeof(double) * U * (size_t)(n_det + 2 * n_amp) * ns);
    double* d_ga = d_gd + U * (size_t)n_det * ns;
    k_distribute_units<<<(unsigned)U, 128, acc_bytes, st>>>(P.slotpart, P.steps, P.unit_steps, S.tape_cap, nC, (int)nred,
                                                          P.tab, prog.dt, ns, n_det, n_amp, d_gd, d_ga);
    PD_CUDA_CHECK(cudaGetLastError());
    ++launches;
    if (g_det && n_det) PD_CUDA_CHECK(copy_out(g_det, d_gd, sizeof(double) * U * n_det * ns, st));
    if (g_amp && n_amp) PD_CUDA_CHECK(copy_out(g_amp, d_ga, sizeof(double) * U * 2 * n_amp * ns, st));
  }
  int ab = 0;
  PD_CUDA_CHECK(copy_d2h(&ab, P.abort_flag, sizeof(int), st));
  PD_CUDA_CHECK(cudaStreamSynchronize(st));
  if (ab) throw Error(PD_ERR_STATE, "small_ket_backward_units: exchange poll timed out");
  (void)max_steps;
  return launches;
}

}  // namespace pd
