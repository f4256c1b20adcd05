#!/usr/bin/env python
"""bench.py -- headline benchmark of the INSITE hot path on B200 (contract: see README / DESIGN.md §6).

    python bench.py --gpus N --steps K --warmup W            # our arm
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm (oracle port of the reference)

Workload (BASELINE.json configs[1]): cancer_sim factual simulation + INSITE population fit at
1M patients x 60 steps per GPU, FP64.  One "step" = one pass of the hot path over the cohort.
metric = executed patient-steps per second (sum over patients of sequence_length-1).

Keys of the JSON line beyond the contract:
  value / ms_per_step   device-resident step on the reference I/O contract: K1 simulate_factual (4 pre-drawn (N,60)
                        draw arrays in, 9 (N,60) arrays + sequence lengths out, + code bytes / per-patient moments for
                        the fit) -> K4 theta_gram_codes -> [all-reduce of 68 doubles] -> K5 STLSQ
  roofline              K1, HBM-bound: algorithmic bytes (6328 B/patient) / CUDA-event time of the launch alone
  roofline_theta_gram   K4, by SURVEY 8d's 1456 B/patient; read_gbs = bytes the lean launch actually reads
  e2e                   host parameters in -> coefficients out, as a stream of cohorts (GeneratedFitPipeline.submit: every
                        step uploads its own pinned inputs and has its own result read back; the next step's upload is
                        queued under the current step): the draws come from the device generator inside the simulator
                        kernel (K1L), chunked H2D overlapped with compute.  e2e.one_step_at_a_time = step_host (a call
                        returns with its result on the host before the next upload starts), e2e.eager the same without
                        the CUDA graph, e2e.h2d_only the copies alone with all ranks uploading at once
  e2e_host_draws        the same through the pre-drawn-array contract (2 GB of draws per step over PCIe)
  device_rng            the generated-draws path with parameters resident (K1L -> K4 -> K5)
  individualisation     config C4: per-patient STLSQ fits/s (K5b) and discovered-ODE rollout steps/s (K6), uniform and
                        irregular sampling, FP64 vs FP32; INSITE BFGS fits/s (K7) with its status histogram
  c3                    config C3: treatment-sequence (K3) and one-step (K2) counterfactual cohorts of the test patients in
                        compact form, their evaluation tau = 1..5 (K9) / one-step (K8) with the population ODE and with one
                        INSITE fit per (patient, t): kernel times, rooflines, the eight RMSEs
  c5                    config C5: the 16M-patient sweep (16M / n_gpus patients per GPU, generated draws, Gram all-reduce)
  cpu_baseline          the oracle's C restatement on one host thread, bounded sample (N=1 only)
`config` is the workload both arms share (the reference arm prints the same dict); `details` holds this arm's choices.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
import warnings

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
warnings.filterwarnings('ignore')

METRIC = "patient-steps/sec (cancer_sim factual + INSITE population fit)"
UNIT = "patient-steps/s"
K1_BYTES_PER_PATIENT = lambda T: 4 * T * 8 + 9 * T * 8 + 10 * 8 + 8      # 6328 B at T=60
K4_BYTES_PER_PATIENT = lambda T: 3 * T * 8 + 8 + 8                         # 1456 B at T=60
K1_KERNELS = {"pitched": "sim_factual_ws<32,1,11,0>",   # csrc/sim_factual_ws.cuh, variant 12 (128-byte-aligned rows)
              "dense": "sim_factual_ws<32,2,6,0>"}      # variant 10 (dense 480-byte rows)
K1_SIDE_KERNELS = {"pitched": "sim_factual_ws<32,1,11,0,side>", "dense": "sim_factual_ws<32,2,6,0,side>"}   # + lean-fit side outputs
K1L_KERNEL = "sim_factual_rng2_kernel<2,4>"             # csrc/sim_factual_rng.cuh: draws generated in registers
K1_SIDE_BYTES_PER_PATIENT = lambda T: ((T + 15) // 16) * 16 + 6 * 8    # lean fit: code bytes + six moment sums written
K4_KERNEL = "theta_gram2_kernel<codes>"     # csrc/theta_gram.cu, fed by the simulator's side outputs (lean fit)
K4_LEAN_BYTES_PER_PATIENT = lambda T: T * 8 + ((T + 15) // 16) * 16 + 8 + 8 + 6 * 8   # volume row, code bytes, seq_len, type, moments


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` capture of the
    same kernel at the bench workload (profiles/r1_traffic.json); None when no capture is recorded."""
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return float(json.load(open(p))[kernel]["dram_bytes_per_launch"])
    except Exception:
        return None


def ncu_entry(kernel):
    p = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(p))[kernel]
    except Exception:
        return None


def bind_to_gpu_numa_node(gpu_index):
    """Pin this process to the CPU cores that are local to its GPU (NVML affinity mask) before any pinned host buffer
    is allocated: cudaHostAlloc places pages on the calling thread's NUMA node, and with one rank per GPU the eight
    host->device streams of a node otherwise share one socket's memory and inter-socket links.  Returns the core
    list, or None when NVML / affinity control is unavailable (nothing is changed then)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (int(m) >> b) & 1]
        cpus = [c for c in cpus if c in os.sched_getaffinity(0)]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return cpus
    except Exception:
        pass
    return None


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    def __init__(self, gpu_index):
        self.rows, self.stop, self.idx = [], threading.Event(), gpu_index
        self.t = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self.stop.wait(0.1)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        import numpy as np
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace('.', '').isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def synth_inputs(n, T, seed):
    """Synthetic cohort of the BASELINE shape: parameters from the reference's generator (host mirror),
    random draws from the device Philox generator (throughput mode)."""
    import numpy as np
    import torch
    import b200_insite.cancer_simulation as cs
    from b200_insite import device as dev
    np.random.seed(seed)
    params = cs.generate_params(n, 2.0, 2.0, 15, 0)
    g = torch.Generator(device='cuda')
    g.manual_seed(1234 + seed)
    noise = 0.01 * torch.randn((n, T), generator=g, device='cuda', dtype=torch.float64)
    rest = [torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64) for _ in range(3)]
    block = torch.from_numpy(dev.pack_params(params))
    static = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.float64))
    return params, block, static, [noise] + rest


def cpu_inputs(n_sample, T, seed=0):
    import numpy as np
    from oracle import rng_export as rx
    np.random.seed(seed)
    params = rx.generate_params(n_sample, 2.0, 2.0, 15, 0)
    return params, rx.draw_factual(n_sample, T)


def cpu_reference_arm(n_sample, T, threads, inputs=None, port_check_patients=1000):
    """The reference's algorithm on the host cores, as fast as a CPU restatement gets: C restatement of
    simulate_factual + get_scaling_params moments + C normal equations of the snippet/FD/library data +
    STLSQ/unbias on them (oracle/), patients split over `threads` host threads.  The numpy/sklearn
    restatement that mirrors the reference's python-loop cost structure is timed on a small subsample
    and reported in `detail`.  Returns (patient_steps_per_s, seconds, detail)."""
    import numpy as np
    from concurrent.futures import ThreadPoolExecutor
    from oracle import rng_export as rx, sim_oracle as so, sindy_np as sp
    params, draws = inputs if inputs is not None else cpu_inputs(n_sample, T)
    static = np.asarray(params['patient_types'], dtype=np.float64)
    so.lib()
    t0 = time.perf_counter()
    sim = so.sim_factual(params, T, draws, n_threads=threads)
    t1 = time.perf_counter()
    seq = sim['sequence_lengths'].astype(np.int64)
    mask = np.arange(T)[None, :] < seq[:, None]
    moments = {k: (sim[k][mask].mean(), sim[k][mask].std()) for k in ('cancer_volume', 'chemo_dosage', 'radio_dosage')}
    bounds = np.linspace(0, n_sample, max(1, threads) + 1).astype(np.int64)

    def part(b):
        lo, hi = int(b[0]), int(b[1])
        sub = {k: sim[k][lo:hi] for k in ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths')}
        return so.theta_gram(sub, static[lo:hi])
    with ThreadPoolExecutor(max(1, threads)) as ex:
        parts = list(ex.map(part, zip(bounds[:-1], bounds[1:])))
    G = sum(p[0] for p in parts); b = sum(p[1] for p in parts)
    coefs, _ = so.stlsq_from_gram(G, b)
    t2 = time.perf_counter()
    steps = float((sim['sequence_lengths'] - 1).sum())
    detail = {"sim_s": t1 - t0, "fit_s": t2 - t1, "patients": n_sample, "threads": threads,
              "sim_only_patient_steps_per_s": steps / (t1 - t0)}
    if port_check_patients:
        m = min(port_check_patients, n_sample)
        sub = {k: (v[:m] if hasattr(v, 'shape') and v.shape[:1] == (n_sample,) else v) for k, v in sim.items()}
        t3 = time.perf_counter()
        means, stds = so.scaling_params(sub)
        data, sc = sp.process_data(sub, means, stds)
        sp.fit_population(data, sc)
        t4 = time.perf_counter()
        detail["numpy_port_fit_s_per_1000_patients"] = (t4 - t3) * 1000.0 / m
    return steps / (t2 - t0), t2 - t0, detail


LOGGED_COEFS = [[-0.05601456082026624, -0.11598756834077, -0.07958279124512227, 0.07326347275734027],
                [-0.5517350343589641, -0.8761667536689084, -0.053397817822270766, -0.035996455669168224],
                [-3.649800303098579, -0.8626472911638889, 1.1157911611997717, -0.6373790072514276],
                [-1.6336216074116419, -3.49858670473956, -3.584018882175004, 0.06151172618047967]]


def workload_config(n, world, T):
    """The workload both arms run (BASELINE.json configs[1]); the reference arm prints the same dict."""
    return {"workload": "cancer_sim factual + INSITE population fit, 1M patients x 60 steps per GPU (FP64)",
            "patients_per_gpu": n, "patients_total": n * world, "seq_length": T, "gamma": 2.0,
            "patient_steps": "executed = sum(sequence_length-1)"}


def _median_ms(fn, reps=5, warm=1):
    import numpy as np
    import torch
    out = None
    for _ in range(warm):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        out = None
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); out = fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts)), out


def block_c4(gen, T, peak):
    """Config C4 on the generated factual cohort of this rank: individualised per-patient fits and discovered-ODE
    rollouts, uniform and irregular sampling, FP64 vs FP32."""
    import numpy as np
    import torch
    from b200_insite import device as dev
    n = gen.n
    xv = gen.volume.contiguous()
    cd = gen.codes[:, :T].contiguous()
    fit_len = gen.sequence_lengths.to(torch.int32)
    prior = gen.coefs.contiguous()
    x32 = xv.to(torch.float32).contiguous()
    out = {"what": "per-patient ridge-to-prior STLSQ fits (K5b, 16 coefficients per patient) on the generated cohort and "
                   "59-step discovered-ODE rollouts with the per-patient coefficients (K6, 5 Euler sub-steps per interval)",
           "patients": n}
    x0 = xv[:, 0].contiguous(); cd1 = cd[:, :T - 1].contiguous()
    g = torch.Generator(device='cuda'); g.manual_seed(11)
    grids = {"uniform": None,
             "irregular": dev.STANDARD_DT * (0.3 + 2.2 * torch.rand((n, T), generator=g, device='cuda', dtype=torch.float64))}
    for name, dts in grids.items():
        d1 = None if dts is None else dts[:, :T - 1].contiguous()
        ms_fit, pc = _median_ms(lambda: dev.stlsq_batched(xv, cd, fit_len, gen.static, prior, 1e4, dts=dts))
        ms_fit32, pc32 = _median_ms(lambda: dev.stlsq_batched(x32, cd, fit_len, gen.static, prior, 1e4, dts=dts))
        pred = torch.empty((n, T - 1), dtype=torch.float64, device='cuda')
        pred32 = torch.empty_like(pred)
        ms_roll, _ = _median_ms(lambda: dev.ode_rollout(x0, gen.static, cd1, pc, drop_below=-1.0, dts=d1, out=pred))
        ms_roll32, _ = _median_ms(lambda: dev.ode_rollout(x0, gen.static, cd1, pc, drop_below=-1.0, dts=d1, out=pred32, fp32=True))
        scale = pc.abs().amax(dim=(1, 2), keepdim=True)
        fit_bytes = n * (T * 8 + T + 4 + 8 + 128 + (0 if dts is None else T * 8))
        roll_bytes = n * (128 + 16 + (T - 1) + (T - 1) * 8 + (0 if dts is None else (T - 1) * 8))
        out[name] = {
            "fits_per_s": n / (ms_fit / 1e3), "fit_ms": ms_fit,
            "fits_per_s_f32_storage": n / (ms_fit32 / 1e3), "fit_f32_storage_ms": ms_fit32,
            "fit_f32_max_rel_dev_vs_f64": float(((pc32 - pc).abs() / scale).max().item()),
            "fit_roofline": {"bound": "fp64 issue / latency (4 Cholesky solves per row)", "bytes_per_launch": fit_bytes,
                             "achieved": fit_bytes / (ms_fit / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
                             "frac": fit_bytes / (ms_fit / 1e3) / 1e9 / peak, "ncu": "profiles/r2_k5b_stlsq_batched_ncu.txt"},
            "rollout_patient_steps_per_s": n * (T - 1) / (ms_roll / 1e3), "rollout_ms": ms_roll,
            "rollout_f32_ms": ms_roll32,
            "rollout_f32_max_rel_dev_vs_f64": float(((pred32 - pred).abs() / pred.abs().clamp_min(1e-3 * float(pred.abs().max()))).max().item()),
            "rollout_roofline": {"bound": "fp64 pipe (2360 FP64 operations per row in the reference's evaluation order)",
                                 "bytes_per_launch": roll_bytes, "achieved": roll_bytes / (ms_roll / 1e3) / 1e9,
                                 "peak": peak, "unit": "GB/s", "frac": roll_bytes / (ms_roll / 1e3) / 1e9 / peak,
                                 "ncu": "profiles/r2_k6_ode_rollout_ncu.txt"}}
        del pc, pc32, pred, pred32
    nb = min(n, 100_000)
    sub = lambda a: a[:nb].contiguous()
    ms7, (c7, status, fval) = _median_ms(lambda: dev.insite_bfgs(sub(xv), sub(cd), sub(fit_len), 1, sub(gen.static), prior, 10.0),
                                         reps=2)
    stn = status.cpu().numpy().astype(np.int64)
    out["insite_bfgs"] = {"kernel": "insite_bfgs_kernel (K7, 16 lanes per row, BFGS over 16 coefficients, FP64; jax line-search "
                                    "semantics, gtol 1e-5)", "rows": nb,
                          "ms": ms7, "fits_per_s": nb / (ms7 / 1e3),
                          "status_hist": {"converged": int(((stn & 255) == 0).sum()), "max_iter": int(((stn & 255) == 1).sum()),
                                          "zoom_failed": int(((stn & 255) == 3).sum()),
                                          "line_search_maxiter": int(((stn & 255) == 5).sum()),
                                          "kept_theta0": int(((stn & 255) == 6).sum()) + int(((stn & 255) == 4).sum()),
                                          "skipped": int((stn < 0).sum())},
                          "mean_iterations": float((stn[stn >= 0] >> 8).mean()),
                          "bound": "FP64 ALU / dependent-issue latency (16 forward sensitivities per Euler sub-step)",
                          "ncu": "profiles/r2_k7_insite_bfgs_ncu.txt"}
    # the two fit options of SURVEY 8(f) F2 on the same cohort: degree-4 library (R factors by tall-skinny QR + SVD-based
    # STLSQ + polynomial rollout) and the smoothing pass of use_smoothed_finite_difference
    try:
        chemo = (cd & 1).to(torch.float64); radio = ((cd >> 1) & 1).to(torch.float64)
        ms_q, rf = _median_ms(lambda: dev.poly_tsqr(xv, chemo, radio, gen.sequence_lengths, gen.static), reps=3)
        ms_s, (pcoef, psup) = _median_ms(lambda: dev.poly_stlsq(rf), reps=3)
        ms_r, _ = _median_ms(lambda: dev.poly_rollout(x0, gen.static, cd1, pcoef), reps=3)
        ms_sm, _ = _median_ms(lambda: dev.smooth_snippets(xv, chemo, radio, gen.sequence_lengths), reps=3)
        rows_q = float(rf[4 * 256:].sum().item())
        out["fit_options"] = {
            "degree4_library": {"kernels": "poly_tsqr + poly_stlsq + poly_rollout (K4p / K5p / K6p)", "sample_rows": rows_q,
                                "tsqr_ms": ms_q, "sample_rows_per_s": rows_q / (ms_q / 1e3), "stlsq_ms": ms_s,
                                "rollout_ms": ms_r, "support_sizes": psup.sum(1).tolist(),
                                "bound": "FP64 issue / latency (Householder folds of 32 sample rows)",
                                "ncu": "profiles/r2_k4p_poly_tsqr_ncu.txt"},
            "smoothed_finite_difference": {"kernel": "smooth_snippets", "ms": ms_sm,
                                           "achieved_gb_per_s": n * T * 8 * 4 / (ms_sm / 1e3) / 1e9}}
        del chemo, radio
    except Exception as e:      # noqa: BLE001
        out["fit_options"] = {"error": repr(e)}
    return out


def block_c3(params, block_dev, static_dev, n, T, H, rank, world, peak, insite_patients):
    """Config C3: counterfactual cohorts of this rank's patients in compact form + their evaluation.
    world > 1: the cohort is n * world patients; every rank simulates the global source prefix itself and then its own
    shard in one launch (counterfactual.sim_cf_shard), evaluates it, and the error sums are all-reduced."""
    import numpy as np
    import torch
    import torch.distributed as dist
    from b200_insite import device as dev
    from b200_insite import counterfactual as cfm
    from b200_insite import compact_eval as ce
    coefs = dev.to_device(np.array(LOGGED_COEFS))
    n_total = n * world
    out = {"what": "test-patient counterfactual cohorts in compact per-patient form (the reference's dense layout would be "
                   f"{n_total * 562 * 65 * 8 * 3 / 1e12:.2f} TB): K3 treatment sequences tau=1..{H} / K2 one-step, draws from "
                   "the device generator keyed by the global patient index; evaluation with the population ODE of the "
                   "reference's log line (K9 / K8) and with one INSITE fit per (patient, t)",
           "patients_total": n_total, "patients_per_gpu": n}
    prefix_blocks = None
    if world > 1:
        # parameters of the global source prefix = the first patients of rank 0
        P = min(n, max(64, n_total // 150))
        pre = block_dev[:, :P].contiguous()
        dist.broadcast(pre, src=0)
        prefix_blocks = (pre, P)

    def make_generator(kind):
        """Draws (device generator, keyed by the global patient index) are made once; the returned callable is the timed
        part: K3 / K2 over this rank's patients (world > 1: source prefix + shard)."""
        extra = H if kind == 'seq' else 0
        dr = cfm.generated_draws(n, T, extra, seed=77, patient_base=rank * n)
        if world == 1:
            if kind == 'seq':
                return lambda: cfm.sim_cf_treatment_seq(block_dev, *dr, T, H)
            return lambda: cfm.sim_cf_one_step(block_dev, *dr, T)
        pre, P = prefix_blocks
        pdr = cfm.generated_draws(P, T, extra, seed=77, patient_base=0)
        return lambda: cfm.sim_cf_shard('treatment_seq' if kind == 'seq' else 'one_step', T, H, n_total, rank * n,
                                        (block_dev,) + tuple(dr), (pre,) + tuple(pdr))[0]

    def allmax(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    def allsum(v):
        if world == 1:
            return v
        t = torch.tensor([float(v)], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t[0])

    for kind in ('seq', 'one'):
        ms_gen, coh = _median_ms(make_generator(kind), reps=3)
        ms_gen = allmax(ms_gen)
        rows = allsum(int(coh.total_rows) if world == 1 else int(coh.n_rows.sum().item()))
        steps = allsum(float(coh.n_steps.double().sum().item()))
        W = T - 1
        if kind == 'seq':
            wbytes = n * (W * 2 * H * H * 8 + T * 9 + W * 2 + 16)
            rbytes = n * ((T + H + 3 * T) * 8 + 80)
            ebytes = n * (W * 2 * H * H * 8 + T + W * 2 + 8 + 12)
            name, ename = "cf_seq_factual_kernel + cf_seq_project_kernel (K3)", "cf_eval_seq_kernel (K9)"
            prof, eprof = "profiles/r2_k3_treatment_seq_ncu.txt", "profiles/r2_k9_cf_eval_seq_ncu.txt"
        else:
            wbytes = n * (W * 4 * 8 + T * 9 + 16)
            rbytes = n * (4 * T * 8 + 80)
            ebytes = n * (W * 4 * 8 + T * 9 + 12)
            name, ename = "cf_one_step_kernel (K2)", "cf_eval_one_step_kernel (K8)"
            prof, eprof = "profiles/r2_k2_one_step_ncu.txt", "profiles/r2_k8_cf_eval_one_step_ncu.txt"
        ms_ev, sums = _median_ms(lambda: ce.evaluate(coh, static_dev, coefs, 1e-3))
        ms_ev = allmax(ms_ev)
        from b200_insite.cohort import allreduce_stats
        allreduce_stats(sums)
        sm = sums.cpu().numpy()
        blk = {"generator": {"kernel": name, "ms": ms_gen, "reference_rows": rows, "rows_per_s": rows / (ms_gen / 1e3),
                             "rollout_steps_per_s": (steps + (H if kind == 'seq' else 1) * rows) / (ms_gen / 1e3),
                             "levels": int(coh.levels),
                             "roofline": {"bound": "hbm (writes) / fp64 log", "bytes_written_per_launch": wbytes,
                                          "bytes_read_per_launch": rbytes,
                                          "achieved": (wbytes + rbytes) * world / (ms_gen / 1e3) / 1e9, "peak": peak * world,
                                          "unit": "GB/s", "frac": (wbytes + rbytes) / (ms_gen / 1e3) / 1e9 / peak, "ncu": prof}},
               "evaluation_population": {"kernel": ename, "ms": ms_ev, "rows_per_s": rows / (ms_ev / 1e3),
                                         "roofline": {"bound": "hbm", "bytes_read_per_launch": ebytes,
                                                      "achieved": ebytes * world / (ms_ev / 1e3) / 1e9, "peak": peak * world,
                                                      "unit": "GB/s", "frac": ebytes / (ms_ev / 1e3) / 1e9 / peak, "ncu": eprof}}}
        if kind == 'seq':
            blk["rmse_tau_1_to_H_percent"] = [float(v) for v in ce.n_step_rmses(sm, H, dev.TUMOUR_DEATH_THRESHOLD)]
        else:
            o, a, l = ce.one_step_rmses(sm, W, dev.TUMOUR_DEATH_THRESHOLD)
            blk["rmse_orig_all_last_percent"] = [float(o), float(a), float(l)]
        # INSITE: one BFGS fit per (patient, t) (lam = 10), then the same evaluation with the per-step coefficients
        ni = min(n, insite_patients)
        if ni > 0:
            sub = cfm.CompactCohort(coh.kind, ni, T, coh.H, coh.factual[:ni].contiguous(), coh.codes[:ni].contiguous(),
                                    coh.cf[:ni].contiguous(), None if coh.valid is None else coh.valid[:ni].contiguous(),
                                    coh.n_steps[:ni].contiguous(), coh.n_rows[:ni].contiguous(), None, 0, 0)
            st_sub = static_dev[:ni].contiguous()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            pc, diag = ce.individualise(sub, st_sub, coefs, estimator='bfgs_rollout', lam=10.0)   # jax semantics + fallback
            a1.record()
            ms_e2, s2 = _median_ms(lambda: ce.evaluate(sub, st_sub, pc, -1.0), reps=3)
            ms_fit = allmax(a0.elapsed_time(a1))
            allreduce_stats(s2)
            s2 = s2.cpu().numpy()
            stn = diag['status'].cpu().numpy().astype(np.int64)
            fits = allsum(int((stn >= 0).sum()))
            ins = {"patients_per_gpu": ni, "fits": fits, "fit_ms": ms_fit, "fits_per_s": fits / (ms_fit / 1e3),
                   "reference_rows_covered": allsum(int(sub.n_rows.sum().item())), "evaluation_ms": allmax(ms_e2),
                   "optimiser": "jax BFGS semantics (gtol 1e-5, zoom failure -> population coefficients, sindy.py:628-631)",
                   "zoom_failed_frac": float(((stn[stn >= 0] & 255) == 3).mean()) if (stn >= 0).any() else 0.0}
            if kind == 'seq':
                ins["rmse_tau_1_to_H_percent"] = [float(v) for v in ce.n_step_rmses(s2, H, dev.TUMOUR_DEATH_THRESHOLD)]
            else:
                o, a, l = ce.one_step_rmses(s2, W, dev.TUMOUR_DEATH_THRESHOLD)
                ins["rmse_orig_all_last_percent"] = [float(o), float(a), float(l)]
            blk["insite"] = ins
            del pc, diag, sub
        out["treatment_seq" if kind == 'seq' else "one_step"] = blk
        del coh
        torch.cuda.empty_cache()
    return out


def block_c5(T, rank, world, total=16_000_000):
    """Config C5: the 16M-patient sweep, 16M / world patients per GPU: generated draws (K1L), lean fit, all-reduce of
    the 68 statistics, STLSQ.  Strong scaling over the GPU counts of the scaling run."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import b200_insite.cancer_simulation as cs
    from b200_insite import device as dev
    from b200_insite.cohort import GeneratedFitPipeline
    n = total // world
    gen = GeneratedFitPipeline(n, T, seed=4321, patient_base=rank * n, chunks=8)
    np.random.seed(100 + rank)
    m = min(n, 1_000_000)                       # parameters: 1M generated on the host, tiled to the shard size
    p = cs.generate_params(m, 2.0, 2.0, 15, 0)
    blk = torch.from_numpy(dev.pack_params(p)).cuda()
    st = torch.from_numpy(np.asarray(p['patient_types'], dtype=np.float64)).cuda()
    reps = -(-n // m)
    gen.params.copy_(blk.repeat(1, reps)[:, :n]); gen.static.copy_(st.repeat(reps)[:n])
    del blk, st
    for _ in range(2):
        gen.step_device()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 5
    a.record()
    for _ in range(k):
        gen.step_device()
    b.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / k
    t = torch.tensor([ms, gen.executed_steps()], dtype=torch.float64, device='cuda')
    if world > 1:
        tm = t.clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t.clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, steps = float(tm[0]), float(ts[1])
    else:
        steps = float(t[1])
    out = {"what": "16M-patient sweep (BASELINE configs[4]): patients sharded over the GPUs, draws generated in the simulator "
                   "kernel (K1L), lean fit, one all-reduce of 68 doubles, STLSQ; parameters resident",
           "patients_total": n * world, "patients_per_gpu": n, "ms_per_step": ms, "value": steps / (ms / 1e3),
           "unit": UNIT, "scaling": "strong (fixed 16M patients)", "population_coefs": gen.coefs.cpu().numpy().tolist()}
    del gen
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = args.ref_patients or args.patients
    vals, detail = [], None
    inputs = cpu_inputs(n_sample, args.seq_length)
    for i in range(args.warmup + args.steps):
        v, sec, detail = cpu_reference_arm(n_sample, args.seq_length, cores, inputs=inputs,
                                           port_check_patients=1000 if i == 0 else 0)
        if i >= args.warmup:
            vals.append((v, sec))
    value = sum(v for v, _ in vals) / len(vals)
    ms = 1e3 * sum(s for _, s in vals) / len(vals)
    sample = (f"{n_sample} patients x {args.seq_length} steps per step: C restatement of simulate_factual + scaling "
              f"moments + C normal equations + STLSQ on {cores} host threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.patients, args.gpus, args.seq_length),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "detail": detail},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from b200_insite import device as dev
    from b200_insite.cohort import FactualFitPipeline, GeneratedFitPipeline

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev.require_cuda()
    n, T = args.patients, args.seq_length        # per GPU (weak scaling)
    params, block, static, draws = synth_inputs(n, T, seed=rank)
    pitch = dev.aligned_pitch(T) if args.layout == "pitched" else T
    pipe = FactualFitPipeline(n, T, variant=args.variant, fused=args.fused, pitch=pitch, lean_fit=bool(args.lean_fit))
    pipe.params.copy_(block.cuda()); pipe.static.copy_(static.cuda())
    for d, s in zip(pipe.draws, draws):
        d.copy_(s)
    del draws

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ("value") ----------------------------------------------------
    for _ in range(args.warmup):
        pipe.step_device()
    barrier()
    k1_ms = []
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    clocks = ClockSampler(local_rank)   # sampled from here to the end of the end-to-end loop (all GPU-busy)
    clocks.t.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        pipe.step_device()
    e1.record()
    barrier()
    total_ms = e0.elapsed_time(e1)
    # dominant kernel alone (same buffers, inputs > L2): CUDA events around the single launch
    def launch_k1():
        if pipe.lean_fit:    # the launch the pipeline uses: simulator + its side outputs for the lean fit
            dev.sim_factual_side(pipe.params, *pipe.draws, T, pipe.consts, out=pipe.out, codes=pipe.codes,
                                 patient_moments=pipe.patient_moments, variant=args.variant)
        else:
            dev.sim_factual(pipe.params, *pipe.draws, T, pipe.consts, out=pipe.out, variant=args.variant,
                            fused_static=pipe.static if args.fused else None)
    for a, b in ev:
        a.record()
        launch_k1()
        b.record()
    torch.cuda.synchronize()
    k1_ms = [a.elapsed_time(b) for a, b in ev]
    # second kernel of the step (population statistics), same way
    for a, b in ev:
        a.record()
        if pipe.lean_fit:
            dev.theta_gram_codes(pipe.out['cancer_volume'], pipe.codes, pipe.out['sequence_lengths'], pipe.static,
                                 pipe.patient_moments)
        else:
            dev.theta_gram(pipe.out['cancer_volume'], pipe.out['chemo_application'], pipe.out['radio_application'],
                           pipe.out['sequence_lengths'], pipe.static, pipe.out['chemo_dosage'], pipe.out['radio_dosage'])
        b.record()
    torch.cuda.synchronize()
    k4_ms = [a.elapsed_time(b) for a, b in ev]
    # the simulator kernel on the reference's dense (N,T) rows, for comparison with the pitched device layout
    k1_dense_ms = None
    if pitch != T and rank == 0:
        dd = [dev.alloc_rows(n, T) for _ in range(4)]
        for d, s_ in zip(dd, pipe.draws):
            d.copy_(s_)
        od = {k_: dev.alloc_rows(n, T) for k_ in dev.FACTUAL_OUT_KEYS}
        od['sequence_lengths'] = torch.empty((n,), dtype=torch.float64, device='cuda')
        for _ in range(2):
            dev.sim_factual(pipe.params, *dd, T, pipe.consts, out=od, variant=args.variant)
        for a, b in ev:
            a.record()
            dev.sim_factual(pipe.params, *dd, T, pipe.consts, out=od, variant=args.variant)
            b.record()
        torch.cuda.synchronize()
        k1_dense_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        same = all(torch.equal(od[k_], pipe.out[k_]) for k_ in od)
        del dd, od
        assert same, "pitched and dense layouts must give bit-identical outputs"
    steps_exec = pipe.executed_steps()
    t = torch.tensor([total_ms, steps_exec], dtype=torch.float64, device='cuda')
    if world > 1:
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        total_ms, steps_exec_all = float(tmax[0]), float(tsum[1])
    else:
        steps_exec_all = steps_exec
    ms_per_step = total_ms / args.steps
    value = steps_exec_all / (ms_per_step / 1e3)

    # ---- end to end through host buffers -----------------------------------------------------------
    pin = lambda x: x.cpu().pin_memory()
    h_block, h_static = pin(block), pin(static)
    h_draws = [pin(d) for d in pipe.draws]
    h_result = torch.empty(32 + dev.STATS_DOUBLES, dtype=torch.float64).pin_memory()
    for _ in range(max(1, min(args.warmup, 2))):
        pipe.step_host(h_block, h_static, h_draws, h_result)
    barrier()
    e2e_steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.step_host(h_block, h_static, h_draws, h_result)
    barrier()
    e2e_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_ms], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = steps_exec_all / (float(te[0]) / 1e3)
    coefs = h_result[:16].numpy().reshape(4, 4).copy()
    del h_draws

    # ---- throughput mode: draws generated inside the simulator kernel (K1L) ------------------------------
    # the reference's simulate_factual draws its random numbers itself (cancer_simulation.py:275-279): the call's
    # inputs are the patient parameters.  Host parameters -> chunked H2D overlapped with K1L -> fit -> D2H result.
    gen = GeneratedFitPipeline(n, T, seed=1234, patient_base=rank * n, chunks=args.rng_chunks)
    gen.params.copy_(pipe.params); gen.static.copy_(pipe.static)
    for _ in range(args.warmup):
        gen.step_device()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        gen.step_device()
    g1.record()
    barrier()
    gen_ms = g0.elapsed_time(g1) / args.steps
    for a, b in ev:
        a.record()
        gen._simulate()
        b.record()
    torch.cuda.synchronize()
    k1l_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
    gen_exec = gen.executed_steps()
    # End to end on the reduced input set: what get_standard_params draws per patient (initial volume, alpha, rho, beta_c,
    # patient type byte) crosses PCIe; beta = alpha / 10, K and the four sigmoid rows are rebuilt on the device from
    # generate_params' own arguments (cancer_simulation.py:83-88, :185, :202) -- nothing is found by scanning the arrays.
    uniform = dev.cohort_scalar_rows(2.0, 2.0) if args.uniform_rows else {}
    h_types = torch.from_numpy(np.asarray(params['patient_types'], dtype=np.uint8)).pin_memory() if args.uniform_rows else None
    if args.uniform_rows:     # outside the timed region: the inputs really are what the reduced contract assumes
        assert np.array_equal(params['beta'], params['alpha'] / 10) and dev.uniform_param_rows(params) == uniform
    def e2e_loop(graph):
        for _ in range(max(1, min(args.warmup, 3))):
            gen.step_host(h_block, h_static, h_result, uniform=uniform, types_u8=h_types, graph=graph)
        barrier()
        t0 = time.perf_counter()
        for _ in range(max(1, args.steps)):
            gen.step_host(h_block, h_static, h_result, uniform=uniform, types_u8=h_types, graph=graph)
        barrier()
        return 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    # eager: ~60 stream operations per step issued from Python; graph: the same step captured once and replayed with one
    # launch (same copies, same kernels, same collective).  This is the one-step-at-a-time number (graph replay unless disabled); the headline e2e is the stream of cohorts below.
    gen_e2e_eager_ms = e2e_loop(False)
    gen_e2e_ms = gen_e2e_eager_ms
    e2e_graph_error = None
    if args.e2e_graph:
        try:
            gen_e2e_ms = e2e_loop(True)
        except Exception as e:      # noqa: BLE001
            e2e_graph_error = repr(e)
    # Stream of cohorts: the next step's upload is queued under the previous step's tail (all-reduce, STLSQ, result copies)
    # -- GeneratedFitPipeline.submit.  Every step still copies its own inputs from pinned host memory and has its own result
    # read back; two sets of pinned input / result buffers alternate, and a buffer is only reused once its copies have passed.
    def e2e_pipelined():
        if h_types is None:
            return None
        blocks = [h_block, h_block.clone().pin_memory()]
        types = [h_types, h_types.clone().pin_memory()]
        results = [h_result, torch.empty_like(h_result).pin_memory()]
        steps_p = max(2, args.steps)
        def run(nsteps):
            pending, last = None, [None, None]
            for s_ in range(nsteps):
                k = s_ & 1
                if last[k] is not None:
                    last[k].inputs_consumed.synchronize()
                st_ = gen.submit(blocks[k], results[k], types[k], uniform=uniform)
                if pending is not None:
                    pending.wait()
                pending, last[k] = st_, st_
            pending.wait()
        run(3)
        barrier()
        t0 = time.perf_counter()
        run(steps_p)
        barrier()
        return 1e3 * (time.perf_counter() - t0) / steps_p
    gen_e2e_pipe_ms, e2e_pipe_error = None, None
    if args.e2e_pipelined:
        try:
            gen_e2e_pipe_ms = e2e_pipelined()
        except Exception as e:      # noqa: BLE001
            e2e_pipe_error = repr(e)
    # What the host->device path alone sustains with every rank copying at once (no kernels): the floor of the end-to-end
    # step when the ranks' uploads share PCIe switches / host memory.  Same bytes, same chunking, same pinned buffers.
    def h2d_probe():
        rows = [r for r in range(10) if r not in (uniform or {})]
        if h_types is not None:
            rows = [r for r in rows if r != 3]           # beta is rebuilt on the device
        for rep in range(3 + max(1, args.steps)):
            if rep == 3:
                barrier()
                t0 = time.perf_counter()
            for a, b in gen.bounds:
                for r in rows:
                    gen.params[r, a:b].copy_(h_block[r, a:b], non_blocking=True)
                if h_types is not None:
                    gen.types_u8_dev[a:b].copy_(h_types[a:b], non_blocking=True)
                else:
                    gen.static[a:b].copy_(h_static[a:b], non_blocking=True)
            torch.cuda.current_stream().synchronize()
        barrier()
        return 1e3 * (time.perf_counter() - t0) / max(1, args.steps)
    try:
        h2d_only_ms = h2d_probe()
    except Exception:      # noqa: BLE001
        h2d_only_ms = None
    clocks.stop.set(); clocks.t.join(timeout=6)
    gen_coefs = h_result[:16].numpy().reshape(4, 4).copy()
    tg = torch.tensor([gen_ms, gen_e2e_ms, gen_e2e_eager_ms, h2d_only_ms or 0.0, gen_e2e_pipe_ms or 0.0], dtype=torch.float64,
                      device='cuda')
    if world > 1:
        tgs = torch.tensor([gen_exec], dtype=torch.float64, device='cuda'); dist.all_reduce(tgs, op=dist.ReduceOp.SUM)
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gen_exec_all = float(tgs[0])
    else:
        gen_exec_all = gen_exec
    gen_ms, gen_e2e_ms, gen_e2e_eager_ms, h2d_only_ms = float(tg[0]), float(tg[1]), float(tg[2]), float(tg[3])
    gen_e2e_pipe_ms = float(tg[4]) or None
    gen_e2e_serial_ms = gen_e2e_ms
    if gen_e2e_pipe_ms:       # the headline end-to-end number is the stream of cohorts; the one-step-at-a-time numbers stay beside it
        gen_e2e_ms = gen_e2e_pipe_ms

    # ---- the other BASELINE configurations (each block is guarded: a failure there must not cost the headline line) ----
    peak_hbm, _ = measured_peak_hbm()
    indiv = c3 = c5 = None
    if rank == 0 and not args.skip_c4:
        try:
            indiv = block_c4(gen, T, peak_hbm)
        except Exception as e:      # noqa: BLE001
            indiv = {"error": repr(e)}
    gen_h2d = gen.h2d_bytes(uniform, reduced=h_types is not None)
    gen_chunks = len(gen.bounds)
    launches = pipe.launches_per_step * args.steps
    lean_fit = bool(pipe.lean_fit)
    pipe_h2d = pipe.h2d_bytes()
    del gen, pipe
    torch.cuda.empty_cache()
    if not args.skip_c3:
        try:
            c3 = block_c3(params, block.cuda(), static.cuda(), n, T, 5, rank, world, peak_hbm, args.insite_patients)
        except Exception as e:      # noqa: BLE001
            c3 = {"error": repr(e)}
            torch.cuda.empty_cache()
    if not args.skip_c5:
        try:
            c5 = block_c5(T, rank, world)
        except Exception as e:      # noqa: BLE001
            c5 = {"error": repr(e)}
            torch.cuda.empty_cache()

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        k1 = float(np.mean(k1_ms))
        k1_name = (K1_SIDE_KERNELS if lean_fit else K1_KERNELS)[args.layout]
        # algorithmic bytes stay the reference I/O contract's; the lean fit's side outputs (112 B/patient) are extra
        achieved = K1_BYTES_PER_PATIENT(T) * n / (k1 / 1e3) / 1e9
        k4 = float(np.mean(k4_ms))
        # K4 is scored on the bytes the launch reads (volume row, code bytes, sequence length, type, six moment sums);
        # SURVEY.md 8(d)'s three-array accounting (1456 B/patient) is kept beside it for comparison only
        k4_read = (K4_LEAN_BYTES_PER_PATIENT(T) if lean_fit else (5 * T * 8 + 16)) * n
        k4_survey = K4_BYTES_PER_PATIENT(T) * n
        achieved4 = k4_read / (k4 / 1e3) / 1e9
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(n, world, T),
                "details": {"nominal_patient_steps_per_s": n * world * (T - 1) / (ms_per_step / 1e3),
                            "cache": "inputs (1.9 GB draws) and outputs (4.3 GB) per step exceed the 126 MB L2",
                            "sim_variant": args.variant, "fused_gram": bool(args.fused), "lean_fit": lean_fit,
                            "layout": (f"(N,{T}) float64 arrays with a row pitch of {pitch} elements ({pitch * 8}-byte rows: "
                                       f"every row starts on a 128-byte line); dense rows are measured beside it in "
                                       f"roofline.dense_rows") if pitch != T else f"dense (N,{T}) float64 rows",
                            "noise": "pre-drawn arrays resident in HBM (reference I/O contract)",
                            "parallelism": f"patients sharded over {world} GPU(s); allreduce of 68 doubles",
                            "host_cores": os.cpu_count(),
                            "host_affinity": (f"rank 0 bound to the {len(numa_cpus)} cores local to its GPU" if numa_cpus
                                              else "default")},
                "clocks": clocks.summary(),
                "e2e": {"value": gen_exec_all / (gen_e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": gen_h2d,
                        "h2d_bytes_per_patient": gen_h2d / n,
                        "rows_rebuilt_on_device": sorted(uniform) + ([3] if h_types is not None else []),
                        "d2h_bytes_per_step": int(h_result.numel() * 8), "ms_per_step": gen_e2e_ms,
                        "launch": ("stream of cohorts (GeneratedFitPipeline.submit, eager stream operations; see stream_of_cohorts)"
                                   if gen_e2e_pipe_ms else
                                   ("one CUDA-graph launch per step (the step's copies, kernels, collective and result copies "
                                    "captured once for these pinned buffers)" if args.e2e_graph and e2e_graph_error is None
                                    else "eager stream operations")),
                        "eager": {"value": gen_exec_all / (gen_e2e_eager_ms / 1e3), "ms_per_step": gen_e2e_eager_ms,
                                  "graph_error": e2e_graph_error},
                        "one_step_at_a_time": {"value": gen_exec_all / (gen_e2e_serial_ms / 1e3), "ms_per_step": gen_e2e_serial_ms,
                                               "what": "step_host: the call returns when its result is on the host, the next "
                                                       "upload starts after that (CUDA-graph replay unless disabled)"},
                        "stream_of_cohorts": {"enabled": bool(gen_e2e_pipe_ms), "error": e2e_pipe_error,
                                              "what": "GeneratedFitPipeline.submit: every step uploads its own pinned inputs and "
                                                      "has its own result read back, but the upload of step s+1 is queued under "
                                                      "the tail of step s (all-reduce, STLSQ, result copies); two sets of pinned "
                                                      "buffers alternate.  This is the e2e value when enabled."},
                        "h2d_only": {"ms_per_step": h2d_only_ms or None,
                                     "gb_per_s_per_gpu": (gen_h2d / (h2d_only_ms * 1e6)) if h2d_only_ms else None,
                                     "what": "the step's host->device copies alone (same pinned buffers, same chunks, no "
                                             "kernels), all ranks copying at once, max over ranks: the floor the shared "
                                             "host->device path puts under the end-to-end step"},
                        "what": "GeneratedFitPipeline.step_host on the reduced input set: pinned host arrays of what "
                                "get_standard_params draws per patient (initial volume, alpha, rho, beta_c as float64, the "
                                "patient type as one byte) -> H2D in "
                                f"{gen_chunks} chunks overlapped with K1L (simulator with the Philox4x32-10 draws generated in "
                                "registers, bit-identical to K1 on the exported draws); beta = alpha / 10, K and the four "
                                "sigmoid rows are rebuilt on the device from generate_params' arguments -> theta_gram_codes "
                                "-> [all-reduce] -> STLSQ -> D2H coefficients/support/statistics.  The cohort is the "
                                "reference's statistically, not bitwise (device generator instead of numpy's sequential "
                                "stream); the bitwise reference contract is e2e_reference_contract"},
                "e2e_reference_contract": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe_h2d,
                                           "d2h_bytes_per_step": int(h_result.numel() * 8), "ms_per_step": float(te[0]),
                                           "what": "FactualFitPipeline.step_host: pinned host params + the four pre-drawn "
                                                   "(N,T) arrays of the K1 contract (the reference's numpy draws) -> H2D -> "
                                                   "K1,K4,K5 -> D2H (PCIe-bound: 2 GB of draws per step)"},
                "device_rng": {"value": gen_exec_all / (gen_ms / 1e3), "unit": UNIT, "ms_per_step": gen_ms,
                               "kernel": K1L_KERNEL, "kernel_ms": k1l_ms,
                               "bound": "FP64 dependent-issue latency / instruction issue (0.68 KB of HBM traffic per "
                                        "patient); ncu: profiles/r1_k1l_gen2_ncu.txt",
                               "ncu": ncu_entry(K1L_KERNEL),
                               "hbm_bytes_per_launch": (80 + T * 8 + ((T + 15) // 16) * 16 + 8 + 48) * n,
                               "what": "GeneratedFitPipeline.step_device: parameters resident, draws generated in the "
                                       "simulator kernel, lean cohort (volume + code bytes + moments) -> fit",
                               "population_coefs": gen_coefs.tolist()},
                "gpu_launches": launches,
                "roofline": {"bound": "hbm", "kernel": k1_name, "achieved": achieved, "peak": peak,
                             "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(k1_name),
                             "peak_source": peak_src, "kernel_ms": k1,
                             "algorithmic_bytes_per_launch": K1_BYTES_PER_PATIENT(T) * n,
                             "share_of_step": k1 / ms_per_step,
                             "side_outputs_bytes_per_launch": K1_SIDE_BYTES_PER_PATIENT(T) * n if lean_fit else 0,
                             "dense_rows": None if k1_dense_ms is None else {
                                 "kernel_ms": k1_dense_ms,
                                 "achieved": K1_BYTES_PER_PATIENT(T) * n / (k1_dense_ms / 1e3) / 1e9,
                                 "frac": K1_BYTES_PER_PATIENT(T) * n / (k1_dense_ms / 1e3) / 1e9 / peak,
                                 "kernel": K1_KERNELS["dense"], "traffic": ncu_traffic(K1_KERNELS["dense"]),
                                 "note": "same arithmetic on the reference's dense 480-byte rows (bit-identical outputs)"}},
                "roofline_theta_gram": {"bound": "fp64 pipe / latency (ncu: FP64 pipe 48 %, issue slots 58 %, "
                                                 "profiles/r1_k4_theta_gram_codes_ncu.txt); HBM figures = bytes the launch reads",
                                        "kernel": K4_KERNEL, "achieved": achieved4, "peak": peak,
                                        "unit": "GB/s", "frac": achieved4 / peak, "traffic": ncu_traffic(K4_KERNEL),
                                        "kernel_ms": k4, "bytes_read_per_launch": k4_read,
                                        "survey_8d_accounting": {"bytes_per_launch": k4_survey,
                                                                 "gbs": k4_survey / (k4 / 1e3) / 1e9,
                                                                 "frac": k4_survey / (k4 / 1e3) / 1e9 / peak,
                                                                 "note": "three-array form (1456 B/patient) the lean launch does "
                                                                         "not move; not a statement about this kernel"},
                                        "fused_step": {"note": "K1 + K4 scored together with K1's bytes (SURVEY 8d rule for a "
                                                               "fused fit)",
                                                       "gbs": K1_BYTES_PER_PATIENT(T) * n / ((k1 + k4) / 1e3) / 1e9,
                                                       "frac": K1_BYTES_PER_PATIENT(T) * n / ((k1 + k4) / 1e3) / 1e9 / peak},
                                        "share_of_step": k4 / ms_per_step},
                "individualisation": indiv,
                "c3": c3,
                "c5": c5,
                "population_coefs": coefs.tolist()}
        if world == 1 and not args.no_cpu_baseline:
            n_cpu = args.ref_patients or 400_000
            v, sec, detail = cpu_reference_arm(n_cpu, T, 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n_cpu} patients x {T} steps: C restatement of "
                                              f"simulate_factual + scaling moments + C normal equations + STLSQ, "
                                              f"1 thread ({sec:.1f} s)", "detail": detail}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--patients", type=int, default=1_000_000, help="patients per GPU")
    ap.add_argument("--seq-length", type=int, default=60)
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--fused", type=int, default=0)
    ap.add_argument("--lean-fit", type=int, default=1, help="simulator side outputs + theta_gram_codes (default) or "
                    "the standalone five-array theta_gram")
    ap.add_argument("--rng-chunks", type=int, default=8, help="H2D/compute overlap chunks of the generated-draws e2e path")
    ap.add_argument("--uniform-rows", type=int, default=1, help="e2e: fill cohort-wide scalar parameter rows on the device "
                    "instead of copying them (0 = copy all ten rows)")
    ap.add_argument("--layout", default="pitched", choices=["pitched", "dense"],
                    help="device-resident (N,T) arrays: rows padded to 128-byte lines, or the reference's dense rows")
    ap.add_argument("--ref-patients", type=int, default=0,
                    help="bounded CPU sample; 0 = the workload's patients per GPU for --impl reference, 400k for the "
                         "single-thread cpu_baseline leg of the b200 arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-graph", type=int, default=1, help="e2e: replay the host step as one CUDA graph (0 = eager only)")
    ap.add_argument("--e2e-pipelined", type=int, default=1,
                    help="e2e: stream of cohorts (the next step's upload overlaps the previous step's tail); 0 = one step at a time")
    ap.add_argument("--skip-c3", action="store_true", help="skip the counterfactual-cohort block (config C3)")
    ap.add_argument("--skip-c4", action="store_true", help="skip the individualisation block (config C4)")
    ap.add_argument("--skip-c5", action="store_true", help="skip the 16M-patient sweep (config C5)")
    ap.add_argument("--insite-patients", type=int, default=1_000_000,
                    help="patients per GPU whose (patient, t) INSITE fits are run in the C3 block (118 fits per patient)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
