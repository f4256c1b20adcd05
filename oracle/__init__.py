"""TEST INFRASTRUCTURE ONLY — CPU oracle for the INSITE hot path.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import, call, link or execute it, and there only as the checker / the
CPU baseline — never as the thing that is measured or shipped.

Contents
--------
ref_loader.py   imports the UNMODIFIED reference simulator from /root/reference (this
                container only; /root/reference does not exist on the GPU box).
sim_c/          plain-C restatement of the reference simulators (A2, A3, A4, A5).
sindy_np.py     numpy restatement of snippeting + FD + library + STLSQ + unbias, the
                Euler rollout, the dataset transforms and the RMSE metrics.
rng_export.py   replays the reference's global-numpy-RNG draw order.
make_golden.py  runs the reference here and writes tests/golden/*.npz.

Parity pinning: see the header of each file; summary in DESIGN.md §3.
"""
