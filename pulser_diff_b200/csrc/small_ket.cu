// Host side of the small-register ket family (kernels: small_ket.cuh, instantiated in
// small_ket_fwd.cu / small_ket_bwd1.cu / small_ket_bwd2.cu so that they compile in parallel).
#include "small_ket.cuh"

namespace pd {
namespace sk {
void launch_forward(int nq, const SkFwd& P, int nC, cudaStream_t st);
void launch_backward(int nq, const SkBwd& P, int nC, cudaStream_t st);
}  // namespace sk
using namespace sk;

// ---- host side ------------------------------------------------------------------------------
struct SmallKetState {
  // device copies of the pulse program
  unsigned long long *d_dm = nullptr, *d_am = nullptr;
  double *d_dv = nullptr, *d_av = nullptr;
  size_t cap_dm = 0, cap_am = 0, cap_dv = 0, cap_av = 0;
  // workspace
  void* ws[16] = {};
  size_t ws_cap[16] = {};
  // stage tape of the most recent recorded forward sweep (slots 14/15) and its identity
  uint64_t tape_gen = 0;
  int tape_steps = -1;
  int tape_cap = 0;
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

// choose elements per thread and cluster size; false if the register does not fit one cluster
// number of CTAs (SK_T real lanes each); false if the register does not fit one cooperative launch
static bool small_shape(size_t L, int batch, int& nC) {
  size_t c = (2 * L + SK_T - 1) / SK_T;
  (void)batch;
  if (c > (size_t)SK_MAXC) return false;
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

static void upload_prog(SmallKetState& S, const Program& prog, const Geometry& g, SkProg& o, cudaStream_t st) {
  auto up = [&](auto*& dptr, size_t& cap, const void* src, size_t bytes) {
    if (cap < bytes) {
      cudaFree(dptr);
      dptr = nullptr;
      PD_CUDA_CHECK(cudaMalloc((void**)&dptr, std::max<size_t>(bytes, 64)));
      cap = bytes;
    }
    if (bytes) PD_CUDA_CHECK(cudaMemcpyAsync(dptr, src, bytes, cudaMemcpyHostToDevice, st));
  };
  static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "mask width");
  up(S.d_dm, S.cap_dm, prog.det_masks.data(), prog.det_masks.size() * 8);
  up(S.d_am, S.cap_am, prog.amp_masks.data(), prog.amp_masks.size() * 8);
  up(S.d_dv, S.cap_dv, prog.det_values.data(), prog.det_values.size() * 8);
  up(S.d_av, S.cap_av, prog.amp_values.data(), prog.amp_values.size() * 8);
  o.nq = prog.nq; o.n_samples = prog.n_samples; o.n_det = prog.n_det(); o.n_amp = prog.n_amp();
  o.dt = prog.dt;
  o.det_masks = S.d_dm; o.det_values = S.d_dv; o.amp_masks = S.d_am; o.amp_values = S.d_av;
  o.diag = g.diag;
}

// Whole forward evolution.  y0 = state at tsave[0], k0 = f(tsave[0], y0), dt0 = initial step.
// Appends every attempt to `records`.  Returns the number of kernel launches.
int small_ket_forward(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                      const pd_options& o, const cplx* y0, const cplx* k0, double dt0,
                      const double* tsave, int n_t, cplx* states, std::vector<pd_step_record>& records,
                      bool want_tape, uint64_t* tape_gen_out, cudaStream_t st) {
  const size_t L = g.dim * (size_t)g.batch;
  int nC;
  if (!small_shape(L, g.batch, nC)) throw Error(PD_ERR_STATE, "small_ket_forward: unsupported shape");
  if (tape_gen_out) *tape_gen_out = 0;
  SkFwd P{};
  upload_prog(S, prog, g, P.prog, st);
  fill_tab(tab, P.tab);
  P.batch = g.batch; P.nC = nC; P.dim = g.dim; P.L = L;
  P.atol = o.atol; P.rtol = o.rtol; P.safety = o.safety_factor; P.minf = o.min_factor; P.maxf = o.max_factor;
  P.max_steps = o.max_steps;
  P.n_replay = o.n_replay;
  const int log_cap = 1 << 15;
  double* d_ts = (double*)S.get(0, sizeof(double) * n_t);
  PD_CUDA_CHECK(cudaMemcpyAsync(d_ts, tsave, sizeof(double) * n_t, cudaMemcpyHostToDevice, st));
  P.tsave = d_ts; P.n_t = n_t;
  P.y_io = (cplx*)S.get(1, sizeof(cplx) * L);
  P.k0_io = (cplx*)S.get(2, sizeof(cplx) * L);
  PD_CUDA_CHECK(cudaMemcpyAsync(P.y_io, y0, sizeof(cplx) * L, cudaMemcpyDeviceToDevice, st));
  PD_CUDA_CHECK(cudaMemcpyAsync(P.k0_io, k0, sizeof(cplx) * L, cudaMemcpyDeviceToDevice, st));
  P.states = states;
  const size_t ys_bytes = sizeof(uint4) * 2 * (2 * L), red_bytes = sizeof(uint4) * 2 * nC * g.batch;
  P.YS = (uint4*)S.get(3, ys_bytes);
  P.red = (uint4*)S.get(4, red_bytes);
  P.log = (pd_step_record*)S.get(5, sizeof(pd_step_record) * log_cap);
  P.log_cap = log_cap;
  P.resume = (SkResume*)S.get(6, sizeof(SkResume));
  P.abort_flag = (int*)S.get(10, 64);
  if (o.n_replay > 0) {
    double* d_rd = (double*)S.get(7, sizeof(double) * o.n_replay);
    unsigned char* d_rc = (unsigned char*)S.get(8, (size_t)o.n_replay);
    PD_CUDA_CHECK(cudaMemcpyAsync(d_rd, o.replay_dt, sizeof(double) * o.n_replay, cudaMemcpyHostToDevice, st));
    PD_CUDA_CHECK(cudaMemcpyAsync(d_rc, o.replay_clipped, (size_t)o.n_replay, cudaMemcpyHostToDevice, st));
    P.replay_dt = d_rd; P.replay_clipped = d_rc;
  }
  if (want_tape) {
    // per accepted step: 6 stage inputs + 6 slopes; sized from the time grid, bounded by memory
    const size_t per_step = 6 * L * sizeof(cplx);
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) fr = (size_t)4 << 30;
    const size_t budget = std::min<size_t>((size_t)3 << 29, fr / 8 + S.ws_cap[14]);
    size_t steps_cap = std::min<size_t>((size_t)16 * n_t + 256, budget / per_step);
    steps_cap = std::max<size_t>(steps_cap, S.ws_cap[14] / per_step);
    if (steps_cap >= 1) {
      P.tapeY = (double*)S.get(14, steps_cap * per_step);
      P.tapeK = (double*)S.get(15, steps_cap * per_step);
      P.tape_cap = (int)std::min<size_t>(steps_cap, (size_t)1 << 30);
    }
    S.tape_steps = -1;
  }
  SkResume r{};
  r.t = tsave[0]; r.dt = dt0; r.error = 1.0; r.cache_dt = dt0; r.cache_err = 1.0;
  r.kk = 0; r.n_acc = 0; r.tape_ok = P.tapeY ? 1 : 0;
  PD_CUDA_CHECK(cudaMemcpyAsync(P.resume, &r, sizeof(r), cudaMemcpyHostToDevice, st));
  int launches = 0;
  std::vector<pd_step_record> chunk;
  for (;;) {
    // the exchange lines carry phase numbers that restart at 1 in every launch
    PD_CUDA_CHECK(cudaMemsetAsync(P.YS, 0, ys_bytes, st));
    PD_CUDA_CHECK(cudaMemsetAsync(P.red, 0, red_bytes, st));
    PD_CUDA_CHECK(cudaMemsetAsync(P.abort_flag, 0, 64, st));
    sk::launch_forward(prog.nq, P, nC, st);
    ++launches;
    PD_CUDA_CHECK(cudaMemcpyAsync(&r, P.resume, sizeof(r), cudaMemcpyDeviceToHost, st));
    PD_CUDA_CHECK(cudaStreamSynchronize(st));
    if (r.n_rec > 0) {
      chunk.resize(r.n_rec);
      PD_CUDA_CHECK(cudaMemcpyAsync(chunk.data(), P.log, sizeof(pd_step_record) * r.n_rec, cudaMemcpyDeviceToHost, st));
      PD_CUDA_CHECK(cudaStreamSynchronize(st));
      records.insert(records.end(), chunk.begin(), chunk.end());
    }
    if (r.status == 1) continue;
    if (r.status == 2) throw Error(PD_ERR_MAX_STEPS, "max_steps reached");
    if (r.status == 3) throw Error(PD_ERR_STATE, "non-finite error norm in DP5 step");
    if (r.status == 4) throw Error(PD_ERR_INVALID, "replay sequence too short");
    if (r.status == 5) throw Error(PD_ERR_STATE, "small_ket_forward: exchange poll timed out");
    break;
  }
  if (P.tapeY && r.tape_ok) {
    S.tape_steps = r.n_acc;
    ++S.tape_gen;
    if (tape_gen_out) *tape_gen_out = S.tape_gen;
  }
  return launches;
}

size_t small_ket_nred(const Program& prog) { return (size_t)prog.n_det() + 2 * (size_t)prog.n_amp() + 1; }

// Whole adjoint sweep over the stage tape recorded by the forward sweep `tape_gen`.
// slot_sums (host, [n_steps*6][nred]) receives per slot the per-term sums (det terms: sum_q gd_q;
// amp terms: sum_q ga_q, sum_q gb_q; last: Re<kbar, k_i>).  Returns the number of launches, or 0
// if that tape is gone (another forward sweep ran on the plan since) or was incomplete: the
// caller then falls back to the stage-by-stage adjoint, which recomputes instead.
int small_ket_backward(SmallKetState& S, const Geometry& g, const Program& prog, const Tableau& tab,
                       const std::vector<double>& tsave, const double* step_t, const double* step_dt,
                       const int* step_interval, const int* step_clipped, int n_steps, uint64_t tape_gen,
                       const cplx* gstates, bool want_coef, double* d_wacc, cplx* lam_out,
                       std::vector<double>& slot_sums, cudaStream_t st) {
  const size_t L = g.dim * (size_t)g.batch;
  int nC;
  if (!small_shape(L, g.batch, nC)) throw Error(PD_ERR_STATE, "small_ket_backward: unsupported shape");
  const int n_t = (int)tsave.size();
  const size_t nred = small_ket_nred(prog);
  if (tape_gen == 0 || tape_gen != S.tape_gen || S.tape_steps != n_steps) return 0;
  if ((size_t)n_steps * 6 * nC * nred * sizeof(double) > ((size_t)1 << 30)) return 0;
  SkBwd P{};
  upload_prog(S, prog, g, P.prog, st);
  fill_tab(tab, P.tab);
  P.batch = g.batch; P.nC = nC; P.dim = g.dim; P.L = L; P.n_t = n_t; P.n_steps = n_steps;
  std::vector<SkStep> hs(std::max(1, n_steps));
  for (int i = 0; i < n_steps; ++i) hs[i] = {step_t[i], step_dt[i], step_interval[i], step_clipped[i]};
  SkStep* d_steps = (SkStep*)S.get(9, sizeof(SkStep) * hs.size());
  PD_CUDA_CHECK(cudaMemcpyAsync(d_steps, hs.data(), sizeof(SkStep) * hs.size(), cudaMemcpyHostToDevice, st));
  P.steps = d_steps;
  P.gstates = gstates;
  P.tapeY = (const double*)S.ws[14];
  P.tapeK = (const double*)S.ws[15];
  P.KB = (uint4*)S.get(3, sizeof(uint4) * 2 * (2 * L));
  PD_CUDA_CHECK(cudaMemsetAsync(P.KB, 0, sizeof(uint4) * 2 * (2 * L), st));
  const size_t n_part = std::max<size_t>(1, (size_t)n_steps * 6 * nC * nred);
  P.slotpart = (double*)S.get(12, sizeof(double) * n_part);
  double* d_we = d_wacc ? (double*)S.get(13, sizeof(double) * L) : nullptr;
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
  slot_sums.assign((size_t)n_steps * 6 * nred, 0.0);
  if (want_coef && n_steps > 0) {
    const size_t n_out = (size_t)n_steps * 6 * nred;
    double* d_out = (double*)S.get(11, sizeof(double) * n_out);
    k_sum_slots<<<(unsigned)((n_out + 255) / 256), 256, 0, st>>>(P.slotpart, d_out, (size_t)n_steps * 6, nC, (int)nred);
    ++launches;
    PD_CUDA_CHECK(cudaMemcpyAsync(slot_sums.data(), d_out, sizeof(double) * n_out, cudaMemcpyDeviceToHost, st));
  }
  PD_CUDA_CHECK(cudaStreamSynchronize(st));
  PD_CUDA_CHECK(cudaGetLastError());
  {
    int ab = 0;
    PD_CUDA_CHECK(cudaMemcpyAsync(&ab, P.abort_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    PD_CUDA_CHECK(cudaStreamSynchronize(st));
    if (ab) throw Error(PD_ERR_STATE, "small_ket_backward: exchange poll timed out");
  }
  return launches;
}

}  // namespace pd
