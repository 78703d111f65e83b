/*
 * pulser_diff_b200 C ABI  --  B200 (sm_100a) evolution engine behind pulser-diff's solver call.
 *
 * This is the drop-in boundary for the ONE hot path of pasqal-io/pulser-diff:
 *
 *     TorchEmulator.run._run_solver                  reference pulser_diff/backend.py:485-529
 *       -> pyqtorch.sesolve(H, psi0, tsave, solver, options)      backend.py:488-494
 *       -> pyqtorch.mesolve(H, rho0, L, tsave, solver, options)   backend.py:502-509
 *     fed by Hamiltonian.build_ham_tensor / H_t       reference pulser_diff/hamiltonian.py:499-548
 *     and differentiated by the autograd tape         reference pulser_diff/derivative.py:40,76
 *
 * The reference hands the solver an opaque closure H(t) -> sparse COO.  Here the same
 * information crosses the boundary as STRUCTURE:
 *
 *   H(t) psi[s] = ( Dint[s] + sum_q d_q(t) r_q(s) ) psi[s]
 *               + sum_q ( bit_q(s) ? g_q(t) : conj(g_q(t)) ) psi[s ^ m_q]
 *
 *   r_q(s) = 1 - bit_q(s)  (bit value 0 = Rydberg; |r> = e0, |g> = e1; hamiltonian.py:296-300)
 *   m_q    = 1 << (N-1-q)  (qubit 0 is the most significant bit; hamiltonian.py:243-268)
 *   Dint[s]= sum_{i<j} U_ij r_i r_j,  U_ij = C6 / r_ij^6   (hamiltonian.py:341-344, 536)
 *   d_q(t) = 2 * sum_{det terms T containing q} interp(det_values[T], t)   (hamiltonian.py:537-540)
 *   g_q(t) =     sum_{amp terms T containing q} interp(amp_values[T], t)   (hamiltonian.py:541-544)
 *   interp(v,t): i1 = max(min(floor(t/dt), n-2), 0); i2 = min(i1+1, n-2);
 *                v[i1] + (v[i2]-v[i1]) * (t - i1*dt) / dt                  (hamiltonian.py:532-542)
 *
 * det_values / amp_values are EXACTLY the reference's coefficient arrays
 * (-0.5*det and 0.5*amp*exp(-i*phase), sub-sampled; hamiltonian.py:421-422, 432).
 *
 * All entry points: plain pointers and sizes, int status (0 = ok), message via
 * pd_last_error().  "dev" pointers are CUDA device pointers on the plan's device,
 * "host" pointers are ordinary host memory.  Complex numbers are interleaved
 * (re, im) doubles (state amplitudes: floats in the complex64 build, see pd_amplitude_bytes).
 * State vectors cross the ABI batch-major: [batch][2^nbits].
 * Work is queued on `stream` (a cudaStream_t passed as void*); calls that return
 * host-side results synchronise that stream before returning.
 */
#ifndef PULSER_DIFF_B200_H
#define PULSER_DIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PD_ABI_VERSION 1

/* status codes */
#define PD_OK 0
#define PD_ERR_INVALID 1   /* bad argument            -> ValueError  */
#define PD_ERR_CUDA 2      /* CUDA runtime failure    -> RuntimeError */
#define PD_ERR_MAX_STEPS 3 /* adaptive solver gave up -> RuntimeError (pyqtorch max_steps) */
#define PD_ERR_STATE 4     /* call order / missing setup -> RuntimeError */

/* state kinds */
#define PD_KET 0     /* psi in C^(2^N)              (sesolve, backend.py:488-494) */
#define PD_DENSITY 1 /* vec(rho) in C^(4^N), row-major rho[r][c] (mesolve, backend.py:502-509) */

/* solver ids, mirroring pyqtorch.utils.SolverType used at backend.py:434,483,487,495 */
#define PD_SOLVER_DP5_SE 0
#define PD_SOLVER_KRYLOV_SE 1
#define PD_SOLVER_DP5_ME 2

typedef struct pd_plan pd_plan; /* one register + pulse program + workspace */
typedef struct pd_tape pd_tape; /* step log + checkpoints of one forward evolution */

/* Solver options: the keys the reference forwards untouched through **options
 * (backend.py:435,493,508); defaults are pyqtorch's. */
typedef struct pd_options {
  double atol;          /* 1e-8  */
  double rtol;          /* 1e-6  */
  int64_t max_steps;    /* 100000 */
  double safety_factor; /* 0.9 */
  double min_factor;    /* 0.2 */
  double max_factor;    /* 5.0 */
  int32_t max_krylov;   /* 80 */
  double exp_tolerance; /* 1e-10 */
  double norm_tolerance;/* 1e-10 */
  /* additions (do not change defaults): */
  int32_t n_replay;     /* >0: force this ACCEPTED-step sequence, every step taken as accepted
                           (shared-step parity protocol) */
  const double* replay_dt;      /* host [n_replay] step sizes (ignored for clipped steps) */
  const uint8_t* replay_clipped;/* host [n_replay] 1 = step lands on the next tsave point */
  int32_t path;         /* kernel family: 0 = auto, 1 = gather, 2 = tiled (18<=N<=23), 3 = small-register
                           cooperative kernels (N<=14), 4 = stream (N>=16), 5 = density tiles (Lindblad,
                           8<=N<=13; opt-in, slower than the gather kernel on B200) */
} pd_options;

typedef struct pd_step_record {
  double t;        /* start time of the attempt (us) */
  double dt;       /* step size used */
  double error;    /* Hairer error norm (accepted iff <= 1) */
  int32_t accepted;
  int32_t clipped; /* dt was cut to land on tsave[interval] */
  int32_t interval;/* index k of the tsave point being integrated towards */
  int32_t _pad;
} pd_step_record;

int pd_abi_version(void);
const char* pd_last_error(void);
void pd_options_default(pd_options* o);

/* ---- plan ------------------------------------------------------------------------------- */
/* kind PD_KET: nbits = n_qubits.  PD_DENSITY: nbits = 2*n_qubits.  `device` = CUDA ordinal. */
int pd_plan_create(pd_plan** out, int32_t n_qubits, int32_t batch, int32_t kind, int32_t device);
int pd_plan_destroy(pd_plan* p);

/* U_ij (host, row-major N*N, upper triangle used).  Builds Dint[2^N] on the device with a
 * kernel.  Replaces 2*int_mat of hamiltonian.py:385-404,536. */
int pd_plan_set_interaction(pd_plan* p, const double* pair_u_host, void* stream);

/* Coefficient terms, the reference's qobj_list split (hamiltonian.py:506-520).
 * det_values: host real [n_det][n_samples]; amp_values: host complex [n_amp][n_samples];
 * masks: bit q set <=> the term's operator acts on qubit q (global channel = all bits).
 * dt = 0.001 / sampling_rate (hamiltonian.py:523). */
int pd_plan_set_terms(pd_plan* p, int32_t n_samples, double dt, int32_t n_det,
                      const uint64_t* det_masks, const double* det_values, int32_t n_amp,
                      const uint64_t* amp_masks, const double* amp_values);

/* Lindblad collapse operators (PD_DENSITY only): n_ops single-qubit 2x2 complex operators,
 * rate already folded in (sqrt(gamma)*op), each applied on EVERY qubit -- the shape produced
 * by _build_collapse_operators (hamiltonian.py:98-143).  ops_host: [n_ops][2][2] complex,
 * (r,g) basis.  n_ops = 0 reproduces backend.py:496-498 (single zero operator). */
int pd_plan_set_collapse(pd_plan* p, int32_t n_ops, const double* ops_host);

/* Kernel family for the next calls on this plan: 0 = auto, 1 = gather, 2 = tiled, 3 = small-register,
 * 4 = stream kernels where the shape allows (used by the family-vs-gather parity tests). */
int pd_plan_set_path(pd_plan* p, int32_t path);

/* ---- single operator applications --------------------------------------------------------- */
/* out = H(t) in   (ket plans; what `H(t) @ psi` does upstream) */
int pd_hpsi(pd_plan* p, void* stream, double t, const void* in_dev, void* out_dev);
/* out = d/dt state: -i H(t) psi  or the Lindblad right-hand side */
int pd_rhs(pd_plan* p, void* stream, double t, const void* in_dev, void* out_dev);

/* ---- evolution ---------------------------------------------------------------------------- */
/* states_dev: [n_t][batch][dim] complex, states[0] = state0.  tsave host [n_t] (us, sorted,
 * tsave[0] is the start time).  If tape_out != NULL a tape is created for pd_evolve_backward
 * (caller destroys it).  n_steps_out/n_rejected_out may be NULL. */
int pd_evolve_forward(pd_plan* p, void* stream, int32_t solver, const pd_options* opt,
                      const void* state0_dev, const double* tsave_host, int32_t n_t,
                      void* states_dev, pd_tape** tape_out);

/* Adjoint sweep over the recorded step sequence (discretise-then-differentiate, like the
 * reference's tape autograd, derivative.py:40,76).  grad_states_dev: cotangent of states
 * [n_t][batch][dim].  Any output may be NULL.  Host outputs: grad_det [n_det][n_samples]
 * real, grad_amp [n_amp][n_samples] complex (dL/dRe + i dL/dIm), grad_pair_u [N*N],
 * grad_tsave [n_t].  grad_state0_dev: [batch][dim]. */
int pd_evolve_backward(pd_plan* p, void* stream, pd_tape* tape, const void* states_dev,
                       const void* grad_states_dev, double* grad_det_host, double* grad_amp_host,
                       double* grad_pair_u_host, double* grad_tsave_host, void* grad_state0_dev);

/* ---- batches of independent parameter sets (BASELINE configs[2]; reference: the user-level loop
 * over QuantumModel parameter sets, docs/gate_optimization.ipynb) ------------------------------
 * n_units evolutions of the SAME register, masks, time grid and options; unit u has its own
 * initial state state0[u] ([batch][dim]) and coefficient tables det_values[u] ([n_det][n_samples])
 * / amp_values[u] ([n_amp][n_samples] complex), laid out unit-major on the host.  states_dev:
 * [n_units][n_t][batch][dim].  Kets with 2*batch*2^N <= 128 run all units in one launch (one CTA
 * per unit); larger units run one after the other.  DP5_SE only. */
int pd_evolve_forward_units(pd_plan* p, void* stream, const pd_options* opt, int32_t n_units,
                            const void* state0_dev, const double* tsave_host, int32_t n_t,
                            const double* det_values_host, const double* amp_values_host,
                            void* states_dev, pd_tape** tape_out);
/* grad_det_host: [n_units][n_det][n_samples], grad_amp_host: [n_units][n_amp][n_samples] complex,
 * grad_state0_dev: [n_units][batch][dim]; any may be NULL. */
int pd_evolve_backward_units(pd_plan* p, void* stream, pd_tape* tape, const void* states_dev,
                             const void* grad_states_dev, const double* det_values_host,
                             const double* amp_values_host, double* grad_det_host, double* grad_amp_host,
                             void* grad_state0_dev);
/* accepted steps of one unit (and its attempted steps), -1 if out of range */
int64_t pd_tape_unit_steps(const pd_tape* t, int32_t unit, int32_t* attempts_out);

int64_t pd_tape_n_records(const pd_tape* t);
int pd_tape_records(const pd_tape* t, pd_step_record* out, int64_t capacity);
int pd_tape_destroy(pd_tape* t);

/* ---- diagonal observables (utils.expect for diagonal O) --------------------- */
/* out_host[n_t]: sum over batch columns of <psi|diag(obs)|psi>  (ket)  or  tr(diag(obs) rho)
 * (density).  obs_dev: real [2^N]. */
int pd_expect_diag(pd_plan* p, void* stream, const void* states_dev, int32_t n_t,
                   const double* obs_dev, double* out_host /* complex [n_t] */);

/* ---- reverse mode of ONE generator application (the autograd node under `H_t(t) @ psi`,
 * reference hamiltonian.py:526-546 recorded on torch's tape; derivative.py:40, 76) ------------ */
/* For k = pd_rhs(t, state) and a cotangent cot on k (torch convention dL/dRe + i dL/dIm):
 *   grad_state_dev [batch*dim] (nullable)  = G(t)^dagger cot,
 *   grad_det_host [n_det*n_samples], grad_amp_host [n_amp*n_samples*2] (nullable): the gradient
 *     w.r.t. the coefficient samples is ADDED (interpolation weights of the reference rule),
 *   grad_pair_host [N*N] (nullable): overwritten with dL/dU_ij (upper triangle),
 *   grad_t_host (nullable): dL/dt through the interpolation,
 *   defer_pair != 0: the per-amplitude interaction weights are accumulated inside the plan
 *     instead (grad_pair_host untouched); pd_pair_gradient_flush reduces them once -- an adjoint
 *     sweep calls this function thousands of times and needs dL/dU_ij only at the end.
 * H(t) psi = i * pd_rhs(t, psi) on ket plans, so the VJP of pd_hpsi is this call with cot' = -i cot. */
int pd_rhs_vjp(pd_plan* p, void* stream, double t, const void* state_dev, const void* cot_dev,
               void* grad_state_dev, double* grad_det_host, double* grad_amp_host,
               double* grad_pair_host, double* grad_t_host, int32_t defer_pair);
/* grad_pair_host [N*N] = dL/dU_ij summed over every deferred pd_rhs_vjp since the last flush
 * (zeros if there was none); clears the accumulator. */
int pd_pair_gradient_flush(pd_plan* p, void* stream, double* grad_pair_host);

/* ---- pieces of one DP5 step for a host-driven stepper (the sharded register, where every
 * generator application contains an exchange step; upstream solver) ------ */
/* out = sum_j w_j * ins[j]  (1 <= n_in <= 8; element-wise, so out may be one of the inputs) */
int pd_lincomb(pd_plan* p, void* stream, void* out_dev, int32_t n_in, const void* const* ins_dev,
               const double* w_host);
/* sumsq_host[batch] = sum over this plan's amplitudes of
 *   |sum_j ew[j] k[j]|^2 / (atol + rtol * max(|y0|, |y1|))^2          (j = 0..6)
 * -- the local share of the DP5 error norm; the caller adds the shares of all ranks and takes
 * sqrt(sum / 2^N).  k_dev[j] may be NULL where ew[j] == 0. */
int pd_dp5_error_sumsq(pd_plan* p, void* stream, const void* const* k_dev, const double* ew_host,
                       const void* y0_dev, const void* y1_dev, double atol, double rtol,
                       double* sumsq_host);

/* ---- sharded register: flips of the qubits that index the GPU (the reference is
 * single-process, SURVEY.md 5.8, so there is no reference site to cite) ----------------------- */
/* out[i] += shift * psi[i] + sum_k coef_k * peer_slices[k][i]  over the plan's 2^N amplitudes
 * (x batch).  peer_slices[k] are device pointers to the partner ranks' slices; they may be
 * peer-mapped (NVLink) memory and are read in place -- no receive buffer.  coef_host: complex
 * [n_peers] (re, im pairs).  The caller orders the launch after the partners' writes. */
int pd_sharded_accumulate(pd_plan* p, void* stream, void* out_dev, const void* psi_dev, double shift,
                          int32_t n_peers, const void* const* peer_slices,
                          const double* coef_host);
/* Same on a range of the slice: every pointer already points at the first amplitude of the range,
 * n_amp amplitudes are updated.  Lets the caller accumulate chunk c while the copy engines are still
 * pulling chunk c + 1 of the partner slices. */
int pd_sharded_accumulate_range(pd_plan* p, void* stream, void* out_dev, const void* psi_dev, double shift,
                                int32_t n_peers, const void* const* peer_slices, const double* coef_host,
                                uint64_t n_amp);

/* ---- measurement hooks (bench.py roofline) ------------------------------------------------- */
/* Average device time (ms, CUDA events on `stream`) of `reps` back-to-back H(t)·psi
 * applications in -> out. */
int pd_bench_hpsi(pd_plan* p, void* stream, double t, int32_t reps, const void* in_dev,
                  void* out_dev, double* ms_per_apply_host);
/* Average device time (ms) of `steps` fixed-size DP5 steps from y_dev (6 generator
 * applications with fused stage combination, y_new, error norm; every step accepted, FSAL). */
int pd_bench_dp5_steps(pd_plan* p, void* stream, double t0, double dt, int32_t steps, void* y_dev,
                       double* ms_per_step_host);

/* ---- introspection ------------------------------------------------------------------------ */
/* number of CUDA kernels this plan has launched since creation (bench.py gpu_launches) */
int64_t pd_plan_launch_count(const pd_plan* p);
/* Bytes this library has copied host->device / device->host (every cudaMemcpyAsync call site of the
 * library is counted) since the process started or since the last call with reset != 0.  bench.py
 * reads it around the timed end-to-end region: e2e.h2d_bytes_per_step / d2h_bytes_per_step. */
int pd_transfer_counters(int64_t* h2d_bytes, int64_t* d2h_bytes, int32_t reset);
/* 1 if this library was built with the CUDA backend, 0 for the host stand-in used by tests */
int pd_is_cuda(void);
/* Bytes per state-vector amplitude of THIS library build: 16 (complex128, libpulser_diff_b200.so -- the
 * precision the reference computes in, backend.py:271, 280) or 8 (complex64, libpulser_diff_b200_c64.so --
 * north_star's optional 1e-5 tier: the same sources compiled with -DPD_C64, same symbols).  In the complex64
 * build every "dev" pointer to amplitudes (states, cotangents, stage vectors, peer slices) holds interleaved
 * (re, im) floats; everything else -- coefficient tables, times, tolerances, gradients w.r.t. samples /
 * pair_u / tsave, expectation values, error norms -- stays double in both builds.  The complex64 build
 * carries the bandwidth-bound kernel families (gather: any ket / density shape; stream: kets of N >= 19;
 * sharded accumulate); registers of N <= 14 are served by the complex128 library (the Python host casts). */
int pd_amplitude_bytes(void);

#ifdef __cplusplus
}
#endif
#endif /* PULSER_DIFF_B200_H */
