"""Device-resident cohort pipeline for the large configurations (1M+ patients per GPU).

One process per GPU; patients are sharded contiguously over ranks (no data-path collective); the only
exchange is a sum-allreduce of the 68 packed population statistics (NCCL over NVLink when a process
group is initialised), after which every rank runs the identical tiny STLSQ and therefore holds
bit-identical population coefficients.

    FactualFitPipeline.step_device()  K1 simulate_factual -> K4 theta_gram (or fused) -> allreduce ->
                                      K5 population STLSQ, inputs already resident in HBM
    FactualFitPipeline.step_host()    the same through host buffers: pinned params + pre-drawn noise
                                      are copied H2D, results (coefficients, support, scaling moments)
                                      are copied back
    GeneratedFitPipeline              throughput mode: the draws come from the device generator inside the
                                      simulator kernel (K1L), so a step's host inputs are the parameters only;
                                      step_host() overlaps their chunked H2D copy with the simulation
Reference path being replaced: SyntheticCancerDataset(mode='factual') -> get_scaling_params ->
SINDY.fit (dataset.py:58-63, cancer_simulation.py:218-375/776-796, sindy.py:145-338).
"""
import numpy as np
import torch

from . import device as dev


def shard_bounds(n_total, rank, world):
    """Contiguous patient range [lo, hi) of `rank`."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(stats, ordered=False):
    """Sum the packed statistics over ranks in place (no-op without a process group).

    ordered=False: one all-reduce (NCCL ring/tree order: fixed for a given world size, so every rank gets
    the same bits, but the bits differ between world sizes at the 1e-16 level).
    ordered=True : all-gather of the per-rank partials + summation in rank order on every rank -- the
    reduction order no longer depends on the collective algorithm."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return stats
    if not ordered:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        return stats
    world = dist.get_world_size()
    parts = [torch.empty_like(stats) for _ in range(world)]
    dist.all_gather(parts, stats)
    acc = parts[0].clone()
    for p in parts[1:]:
        acc += p
    stats.copy_(acc)
    return stats


def pack_stats(G, b, counts, moments=None):
    """(G (4,4,4), b (4,4), counts (4,), moments (8,)) -> the 68-double packed layout of include/b200i.h."""
    out = np.zeros(dev.STATS_DOUBLES)
    for a in range(4):
        k = a * dev.GRAM_PER_TREATMENT
        for i in range(4):
            for j in range(i, 4):
                out[k] = G[a][i][j]
                k += 1
        out[a * dev.GRAM_PER_TREATMENT + 10:a * dev.GRAM_PER_TREATMENT + 14] = b[a]
        out[a * dev.GRAM_PER_TREATMENT + 14] = counts[a]
    if moments is not None:
        out[4 * dev.GRAM_PER_TREATMENT:] = moments
    return out


def chunk_bounds(n, chunks):
    """Column ranges of the chunked upload (the formula of b200i_upload_simulate_rng): at most `chunks` ranges of
    whole 32-patient tiles that cover [0, n)."""
    n, chunks = int(n), max(1, min(int(chunks), n // 32 if n >= 32 else 1))
    step = -(-n // chunks)
    step = -(-step // 32) * 32
    return [(a, min(a + step, n)) for a in range(0, n, step)]


def source_prefix_length(row_offsets, n_total):
    """Number of leading patients whose rows cover row indices [0, n_total): the global source prefix every
    rank must simulate before its own shard of a counterfactual cohort (cross-row window)."""
    off = np.asarray(row_offsets)
    return int(np.searchsorted(off, n_total, side='left'))


class FactualFitPipeline:
    def __init__(self, n_local, T=60, window_size=15, threshold=1e-3, alpha=0.5, max_iter=100, variant=0,
                 fused=False, pitch=None, lean_fit=True):
        """pitch: row pitch (elements) of the device-resident (N,T) arrays; None = dense rows (the reference's
        numpy layout), dev.aligned_pitch(T) = rows padded to 128-byte lines (the faster layout)."""
        dev.require_cuda()
        self.n, self.T = int(n_local), int(T)
        self.consts = dev.sim_consts(window_size, 0)
        self.threshold, self.alpha, self.max_iter = threshold, alpha, max_iter
        self.variant, self.fused = variant, fused
        f64 = dict(dtype=torch.float64, device='cuda')
        self.params = torch.empty((10, self.n), **f64)
        self.static = torch.empty((self.n,), **f64)
        # lean_fit: the simulator kernel also writes one treatment-code byte per step and six per-patient moment sums,
        # and theta_gram_codes finishes the statistics from the volumes + those (0.6 instead of 2.4 GB read)
        self.lean_fit = bool(lean_fit) and not fused
        self.pitch = self.T if pitch is None else int(pitch)
        self.draws = [dev.alloc_rows(self.n, self.T, self.pitch) for _ in range(4)]   # noise, recovery, chemo, radio
        self.out = {k: dev.alloc_rows(self.n, self.T, self.pitch) for k in dev.FACTUAL_OUT_KEYS}
        self.out['sequence_lengths'] = torch.empty((self.n,), **f64)
        if self.lean_fit:
            self.codes = torch.zeros((self.n, ((self.T + 15) // 16) * 16), dtype=torch.uint8, device='cuda')
            self.patient_moments = torch.empty((6, self.n), **f64)
        self.stats = torch.zeros(dev.STATS_DOUBLES, **f64)
        self.coefs = None
        self.support = None
        self.launches_per_step = 2 if fused else 3   # simulate_factual (+ theta_gram) + stlsq_population

    # -- inputs ------------------------------------------------------------------------------------
    def load_host(self, params_block, static, draws, non_blocking=True):
        """H2D copy of one step's inputs from (pinned) host tensors."""
        self.params.copy_(params_block, non_blocking=non_blocking)
        self.static.copy_(static, non_blocking=non_blocking)
        for d, h in zip(self.draws, draws):
            d.copy_(h, non_blocking=non_blocking)

    def h2d_bytes(self):
        return (10 + 1 + 4 * self.T) * self.n * 8

    # -- one pass of the hot path ------------------------------------------------------------------
    def step_device(self):
        if self.lean_fit:
            out, _, _ = dev.sim_factual_side(self.params, *self.draws, self.T, self.consts, out=self.out, codes=self.codes,
                                             patient_moments=self.patient_moments, variant=self.variant)
            stats = dev.theta_gram_codes(out['cancer_volume'], self.codes, out['sequence_lengths'], self.static,
                                         self.patient_moments)
            self.stats.copy_(stats)
            allreduce_stats(self.stats)
            self.coefs, self.support = dev.stlsq_population(self.stats, self.threshold, self.alpha, self.max_iter)
            return self.coefs
        out, stats = dev.sim_factual(self.params, *self.draws, self.T, self.consts, out=self.out,
                                     variant=self.variant, fused_static=self.static if self.fused else None)
        if not self.fused:
            stats = dev.theta_gram(out['cancer_volume'], out['chemo_application'], out['radio_application'],
                                   out['sequence_lengths'], self.static, out['chemo_dosage'], out['radio_dosage'])
        self.stats.copy_(stats)
        allreduce_stats(self.stats)
        self.coefs, self.support = dev.stlsq_population(self.stats, self.threshold, self.alpha, self.max_iter)
        return self.coefs

    def step_host(self, params_block, static, draws, result_host):
        """Host buffers in, host result out: result_host is a pinned (16 + 16 + 68,) float64 tensor that
        receives coefficients, support and the packed statistics."""
        self.load_host(params_block, static, draws)
        self.step_device()
        result_host[:16].copy_(self.coefs.reshape(-1), non_blocking=True)
        result_host[16:32].copy_(self.support.reshape(-1).to(torch.float64), non_blocking=True)
        result_host[32:32 + dev.STATS_DOUBLES].copy_(self.stats, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return result_host

    def executed_steps(self):
        """Executed simulator loop iterations of the last step: sum(sequence_length - 1)."""
        return float((self.out['sequence_lengths'] - 1.0).sum().item())


class HostStep:
    """One submitted step of GeneratedFitPipeline.submit: its pinned result buffer and the two events of the protocol."""

    def __init__(self, result_host, result_ready, inputs_consumed):
        self.result_host, self.result_ready, self.inputs_consumed = result_host, result_ready, inputs_consumed

    def wait(self):
        self.result_ready.synchronize()
        return self.result_host


class GeneratedFitPipeline:
    """Factual cohort + population fit with device-generated draws (throughput mode).

    The reference's simulate_factual draws its random numbers itself (cancer_simulation.py:275-279), so the call's
    inputs are the patient parameters; here the draws come from Philox4x32-10 counted by the global patient index
    inside the simulator kernel (K1L, b200i_sim_factual_rng) and the cohort is kept in its lean device form:
    volume (N,T) float64 + one treatment-code byte per step + sequence lengths + six per-patient moment sums,
    which is exactly what the fit (b200i_theta_gram_codes -> all-reduce -> b200i_stlsq_population) consumes."""

    def __init__(self, n_local, T=60, seed=0, patient_base=0, window_size=15, threshold=1e-3, alpha=0.5, max_iter=100,
                 chunks=8):
        dev.require_cuda()
        self.n, self.T = int(n_local), int(T)
        self.seed, self.patient_base = int(seed), int(patient_base)
        self.consts = dev.sim_consts(window_size, 0)
        self.threshold, self.alpha, self.max_iter = threshold, alpha, max_iter
        f64 = dict(dtype=torch.float64, device='cuda')
        self.params = torch.empty((10, self.n), **f64)
        self.static = torch.empty((self.n,), **f64)
        self.volume = dev.alloc_rows(self.n, self.T, dev.aligned_pitch(self.T))
        self.codes = torch.empty((self.n, ((self.T + 15) // 16) * 16), dtype=torch.uint8, device='cuda')
        self.sequence_lengths = torch.empty((self.n,), **f64)
        self.patient_moments = torch.empty((6, self.n), **f64)
        self.stats = torch.zeros(dev.STATS_DOUBLES, **f64)
        self.coefs = self.support = None
        self.bounds = chunk_bounds(self.n, chunks)                   # what the C side will use
        self.chunks = len(self.bounds)
        self.copy_stream = torch.cuda.Stream()
        self.chunk_ws = dev.chunk_workspaces(self.chunks)
        self.launches_per_step = len(self.bounds) + 2               # K1L per chunk + theta_gram_codes + stlsq_population

    def h2d_bytes(self, uniform=None, reduced=False):
        if reduced:      # four drawn parameter rows + one patient-type byte
            return (10 - len(uniform or {}) - 1) * self.n * 8 + self.n
        return (10 - len(uniform or {}) + 1) * self.n * 8

    def _simulate(self, rows=None):
        dev.sim_factual_rng(self.params, self.T, self.seed, patient_base=self.patient_base, consts=self.consts,
                            volume=self.volume, codes=self.codes, sequence_lengths=self.sequence_lengths,
                            patient_moments=self.patient_moments, rows=rows)

    def _fit(self):
        stats = dev.theta_gram_codes(self.volume, self.codes, self.sequence_lengths, self.static, self.patient_moments)
        self.stats.copy_(stats)
        allreduce_stats(self.stats)
        self.coefs, self.support = dev.stlsq_population(self.stats, self.threshold, self.alpha, self.max_iter)
        return self.coefs

    def step_device(self):
        """Parameters already resident in HBM: one K1L launch over the shard, then the fit."""
        self._simulate()
        return self._fit()

    def _enqueue_host_step(self, params_block, static, result_host, uniform, types_u8, pipelined=False):
        """Everything of step_host but the final synchronisation, on the current stream.
        pipelined (submit): the statistics of consecutive steps alternate between two buffers and the tail of the step --
        all-reduce, STLSQ, result copies -- runs on a stream of its own, so the next step's simulation queues right behind
        this step's and its upload (b200i_upload_simulate_rng_pipelined) under both.  Returns the stream the result copies
        were queued on."""
        stats, tail = self.stats, None
        if pipelined:
            if getattr(self, '_stats_ring', None) is None:
                self._stats_ring = [torch.zeros_like(self.stats) for _ in range(2)]
                self._tail_stream = torch.cuda.Stream()
                self._submitted = 0
            stats = self._stats_ring[self._submitted & 1]
            self._submitted += 1
            tail = self._tail_stream
        if types_u8 is not None:
            if getattr(self, 'types_u8_dev', None) is None:
                self.types_u8_dev = torch.empty((self.n,), dtype=torch.uint8, device='cuda')
            dev.upload_simulate_rng(params_block, None, self.params, self.static, self.T, self.seed, self.patient_base,
                                    self.consts, self.volume, self.codes, self.sequence_lengths, self.patient_moments,
                                    self.chunks, self.copy_stream, chunk_ws=self.chunk_ws, stats_out=stats,
                                    uniform=uniform, derive_beta=True, types_u8_host=types_u8, types_u8_dev=self.types_u8_dev,
                                    pipelined=pipelined)
        else:
            dev.upload_simulate_rng(params_block, static, self.params, self.static, self.T, self.seed, self.patient_base,
                                    self.consts, self.volume, self.codes, self.sequence_lengths, self.patient_moments,
                                    self.chunks, self.copy_stream, chunk_ws=self.chunk_ws, stats_out=stats,
                                    uniform=uniform)
        main = torch.cuda.current_stream()
        if tail is not None:
            ready = torch.cuda.Event()
            ready.record(main)
            tail.wait_event(ready)
        with torch.cuda.stream(tail if tail is not None else main):
            allreduce_stats(stats)
            self.coefs, self.support = dev.stlsq_population(stats, self.threshold, self.alpha, self.max_iter)
            result_host[:16].copy_(self.coefs.reshape(-1), non_blocking=True)
            result_host[16:32].copy_(self.support.reshape(-1).to(torch.float64), non_blocking=True)
            result_host[32:32 + dev.STATS_DOUBLES].copy_(stats, non_blocking=True)
        return tail if tail is not None else main

    def step_host(self, params_block, static, result_host, uniform=None, types_u8=None, graph=False):
        """Pinned host parameters (10,N) + static feature (N,) in, pinned result (16 + 16 + 68,) out.
        types_u8: pinned uint8 (N,) patient types = the REDUCED input set (what get_standard_params draws: initial volume,
        alpha, rho, beta_c, patient type); `static` is then ignored, beta = alpha / 10 and the static feature are rebuilt
        on the device as generate_params builds them: 33 bytes per patient with the five scalar rows in `uniform`.
        uniform: {row: value} of parameter rows that are one scalar for the cohort (device.cohort_scalar_rows): they
        are filled on the device instead of being copied.  The parameter rows of chunk c+1 are copied on a second stream
        while chunk c is being simulated and reduced to its share of the population statistics; the shares are summed in
        chunk order (bits depend on the chunk count only).
        graph=True: the whole step (chunked copies on the copy stream, per-chunk kernels on two compute streams, the
        all-reduce, STLSQ and the device->host copies of the result) is captured once into a CUDA graph for these host
        buffers and replayed: one graph launch instead of ~60 stream operations per step (the copies read the same
        pinned buffers every time, so the caller refreshes their CONTENT between steps)."""
        main = torch.cuda.current_stream()
        if not graph:
            self._enqueue_host_step(params_block, static, result_host, uniform, types_u8)
            main.synchronize()
            return result_host
        key = (params_block.data_ptr(), None if static is None else static.data_ptr(), result_host.data_ptr(),
               None if types_u8 is None else types_u8.data_ptr(), tuple(sorted((uniform or {}).items())))
        if getattr(self, '_graph_key', None) != key:
            self._enqueue_host_step(params_block, static, result_host, uniform, types_u8)   # warm-up outside the capture
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._enqueue_host_step(params_block, static, result_host, uniform, types_u8)
            self._graph, self._graph_key = g, key
        self._graph.replay()
        main.synchronize()
        return result_host

    def submit(self, params_block, result_host, types_u8, uniform=None):
        """step_host for a stream of cohorts: queues the whole step (reduced input set) and returns at once with a
        HostStep.  The upload of this step runs under the previous step's kernels and tail, and its simulation queues right
        behind the previous step's simulation: the tail (all-reduce, STLSQ, result copies) has a stream of its own.  Protocol: keep `params_block` / `types_u8` unchanged until step.inputs_consumed has
        passed, read `result_host` after step.wait().  Results are bit-identical to step_host's."""
        tail = self._enqueue_host_step(params_block, None, result_host, uniform, types_u8, pipelined=True)
        done = torch.cuda.Event(); done.record(tail)
        consumed = torch.cuda.Event(); consumed.record(self.copy_stream)
        return HostStep(result_host, done, consumed)

    def executed_steps(self):
        return float((self.sequence_lengths - 1.0).sum().item())
