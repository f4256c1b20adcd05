"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous patient sharding, the sum
all-reduce of the packed population statistics (plain and rank-ordered) and the counterfactual source
prefix.  Per-shard statistics come from the oracle's C normal equations (tests may use the oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as h


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, ordered, out_dir):
    sys.path.insert(0, h.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_insite import cohort
    from oracle import sim_oracle as so
    params, draws = h.random_cohort(n_total, seed=13)
    lo, hi = cohort.shard_bounds(n_total, rank, world)
    sub_p = {k: (v[lo:hi] if isinstance(v, np.ndarray) else v) for k, v in params.items()}
    sub_d = {k: v[lo:hi] for k, v in draws.items()}
    sim = so.sim_factual(sub_p, 60, sub_d)
    G, b, cnt = so.theta_gram(sim, sub_p['patient_types'])
    stats = torch.from_numpy(cohort.pack_stats(G, b, cnt))
    cohort.allreduce_stats(stats, ordered=ordered)
    np.save(os.path.join(out_dir, f"stats_{int(ordered)}_{rank}.npy"), stats.numpy())
    dist.destroy_process_group()


@pytest.mark.parametrize("ordered", [False, True])
def test_sharded_statistics_allreduce_equals_whole_cohort(tmp_path, ordered):
    from b200_insite import cohort, device as dev
    from oracle import sim_oracle as so
    n_total, world = 1001, 2          # odd: shards of 501 and 500
    mp.spawn(_worker, args=(world, _free_port(), n_total, ordered, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"stats_{int(ordered)}_{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1]), "ranks must hold bit-identical statistics"
    params, draws = h.random_cohort(n_total, seed=13)
    sim = so.sim_factual(params, 60, draws)
    G, b, cnt = so.theta_gram(sim, params['patient_types'])
    u = dev.unpack_stats(got[0])
    assert np.array_equal(u['count'], cnt)
    np.testing.assert_allclose(u['G'], G, rtol=1e-13)
    np.testing.assert_allclose(u['b'], b, rtol=1e-10, atol=1e-8)
    c_ref, s_ref = so.stlsq_from_gram(G, b)
    c_got, s_got = so.stlsq_from_gram(u['G'], u['b'])
    assert np.array_equal(s_ref, s_got)
    np.testing.assert_allclose(c_got, c_ref, rtol=1e-9)


def test_shard_bounds_cover_the_cohort_exactly():
    from b200_insite.cohort import shard_bounds
    for n in (0, 1, 7, 8, 1_000_003):
        for world in (1, 2, 4, 8):
            b = [shard_bounds(n, r, world) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_source_prefix_length():
    from b200_insite.cohort import source_prefix_length
    off = np.array([0, 236, 472, 700, 936])        # rows emitted before each patient
    assert source_prefix_length(off, 1) == 1       # row 0 lives in patient 0
    assert source_prefix_length(off, 236) == 1
    assert source_prefix_length(off, 237) == 2
    assert source_prefix_length(off, 936) == 4


def test_generated_cohort_sharding_is_a_partition_of_the_global_counter_space():
    """GeneratedFitPipeline: rank r simulates global patients [lo_r, hi_r) (patient_base = lo_r), and within a rank the
    upload chunks are whole 32-patient tiles that cover the shard once: every global patient index -- the counter of
    the draw generator -- is simulated exactly once for any world size and chunk count."""
    from b200_insite.cohort import shard_bounds, chunk_bounds
    for n_total in (1, 31, 32, 1000, 4133, 1_000_003):
        for world in (1, 2, 3, 8):
            seen = 0
            for rank in range(world):
                lo, hi = shard_bounds(n_total, rank, world)
                assert lo == seen
                for chunks in (1, 4, 16, 64):
                    b = chunk_bounds(hi - lo, chunks) if hi > lo else []
                    assert len(b) <= chunks
                    assert all(x[1] == y[0] for x, y in zip(b, b[1:]))
                    assert (not b) or (b[0][0] == 0 and b[-1][1] == hi - lo)
                    assert all(a % 32 == 0 for a, _ in b)
                seen = hi
            assert seen == n_total


def _cf_worker(rank, world, port, n_total, kind, out_dir):
    """Sharded counterfactual cohort with the oracle standing in for the GPU kernels: every rank simulates the global
    source prefix (reference verbatim), then its own shard against the prefix's rows (windowed oracle), then the
    ranks exchange their row totals (counterfactual.exchange_row_bases -- the host logic under test)."""
    sys.path.insert(0, h.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_insite import cohort, counterfactual as cf
    from oracle import sim_oracle as so
    T, H = 60, 5
    seq = kind == 'seq'
    params, draws = h.random_cohort(n_total, seed=17, extra=H if seq else 0)
    sub = lambda p, s: {k: (v[s] if isinstance(v, np.ndarray) else v) for k, v in p.items()}
    run = (lambda p, d, **kw: so.sim_cf_treatment_seq(p, T, H, d, **kw)) if seq else \
          (lambda p, d, **kw: so.sim_cf_one_step(p, T, d, **kw))
    P = min(n_total, n_total // cf.rows_per_patient_floor('treatment_seq' if seq else 'one_step', H) + 2)
    pre = run(sub(params, slice(0, P)), {k: v[:P] for k, v in draws.items()})
    assert pre['cancer_volume'].shape[0] >= n_total
    lo, hi = cohort.shard_bounds(n_total, rank, world)
    if lo == 0:   # patient 0 reads the row it is writing itself: the first shard starts with the verbatim semantics
        mine = run(sub(params, slice(0, hi)), {k: v[:hi] for k, v in draws.items()})
    else:
        mine = run(sub(params, slice(lo, hi)), {k: v[lo:hi] for k, v in draws.items()},
                   window_rows=pre['cancer_volume'][lo:hi])
    base, total = cf.exchange_row_bases(mine['cancer_volume'].shape[0])
    np.savez(os.path.join(out_dir, f"cf_{kind}_{rank}.npz"), base=base, total=total, **mine)
    dist.destroy_process_group()


@pytest.mark.parametrize("kind,n_total", [('one', 1301), ('seq', 1203)])
def test_sharded_counterfactual_cohort_equals_whole_cohort(tmp_path, kind, n_total):
    """SURVEY 8(e) exception: shards + global source prefix + one all-gather of row totals = the single-process cohort."""
    from oracle import sim_oracle as so
    world = 2
    mp.spawn(_cf_worker, args=(world, _free_port(), n_total, kind, str(tmp_path)), nprocs=world, join=True)
    params, draws = h.random_cohort(n_total, seed=17, extra=5 if kind == 'seq' else 0)
    ref = so.sim_cf_treatment_seq(params, 60, 5, draws) if kind == 'seq' else so.sim_cf_one_step(params, 60, draws)
    parts = [np.load(tmp_path / f"cf_{kind}_{r}.npz") for r in range(world)]
    R = ref['cancer_volume'].shape[0]
    assert int(parts[0]['base']) == 0 and int(parts[1]['base']) == parts[0]['cancer_volume'].shape[0]
    assert int(parts[0]['total']) == int(parts[1]['total']) == R
    for k in ('cancer_volume', 'chemo_application', 'radio_application', 'sequence_lengths'):
        got = np.concatenate([p[k] for p in parts])
        assert np.array_equal(got, ref[k], equal_nan=True), k


def _eval_worker(rank, world, port, out_dir):
    """Sharded evaluation of a counterfactual test set: every rank holds the error sums of its patients (here from the
    numpy restatement of the dense rows), the sums are all-reduced and every rank derives the same RMSEs."""
    sys.path.insert(0, h.ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from b200_insite import cohort, compact_eval as ce, counterfactual as cf
    data = np.load(os.path.join(out_dir, "sums_in.npz"))
    one = torch.from_numpy(data[f"one_{rank}"].copy())
    seq = torch.from_numpy(data[f"seq_{rank}"].copy())
    cohort.allreduce_stats(one)
    cohort.allreduce_stats(seq)
    base, total = cf.exchange_row_bases(int(data[f"rows_{rank}"]))
    np.savez(os.path.join(out_dir, f"eval_{rank}.npz"), one=one.numpy(), seq=seq.numpy(), base=base, total=total,
             r1=np.array(ce.one_step_rmses(one.numpy(), 59, 1150.0)), r5=ce.n_step_rmses(seq.numpy(), 5, 1150.0))
    dist.destroy_process_group()


def test_sharded_evaluation_sums_and_row_bases(tmp_path):
    """compact_eval on N > 1 ranks: additive error sums (all-reduce), row bases by an exclusive scan of the per-rank row
    totals (counterfactual.exchange_row_bases) -- the two exchanges of a sharded config-C3 evaluation."""
    from b200_insite import compact_eval as ce
    rng = np.random.RandomState(5)
    W, H, world = 59, 5, 2
    parts = {}
    rows = [4 * 1234, 4 * 987]
    for r in range(world):
        cnt = np.sort(rng.randint(1, rows[r], size=W))[::-1].astype(np.float64)
        cnt[0] = rows[r]
        se = rng.rand(W) * cnt * 100.0
        last = rng.rand(W) * 50.0
        parts[f"one_{r}"] = np.concatenate([se, cnt, last, [last.sum(), rows[r]]])
        parts[f"seq_{r}"] = np.concatenate([rng.rand(H) * 1e4, np.full(H, 10.0 * rows[r])])
        parts[f"rows_{r}"] = np.array(rows[r])
    np.savez(tmp_path / "sums_in.npz", **parts)
    mp.spawn(_eval_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"eval_{r}.npz") for r in range(world)]
    tot_one = parts["one_0"] + parts["one_1"]
    tot_seq = parts["seq_0"] + parts["seq_1"]
    for r in range(world):
        assert np.array_equal(got[r]["one"], tot_one) and np.array_equal(got[r]["seq"], tot_seq)
        assert int(got[r]["total"]) == sum(rows) and int(got[r]["base"]) == sum(rows[:r])
        np.testing.assert_allclose(got[r]["r1"], ce.one_step_rmses(tot_one, W, 1150.0), rtol=1e-15)
        np.testing.assert_allclose(got[r]["r5"], ce.n_step_rmses(tot_seq, H, 1150.0), rtol=1e-15)
    # the formulas themselves: time_varying_model.py:247-281 / :298-311 on the summed quantities
    orig, all_, last = ce.one_step_rmses(tot_one, W, 1150.0)
    se, cnt = tot_one[:W], tot_one[W:2 * W]
    assert np.isclose(all_, np.sqrt(se.sum() / cnt.sum()) / 1150.0 * 100)
    assert np.isclose(orig, np.sqrt((se / cnt).mean()) / 1150.0 * 100)
    assert np.isclose(last, np.sqrt(tot_one[3 * W] / tot_one[3 * W + 1]) / 1150.0 * 100)
