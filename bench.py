#!/usr/bin/env python
"""bench.py -- headline benchmark of the LIBXSMM hot path on B200 (contract: one JSON line on stdout).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload brgemm|fsspmdm|bcsc|sweep]

Workload (BASELINE.json configs[1]): batched BRGEMM bf16 x bf16 -> f32, m=n=k=64, br=8 (stride mode),
batch = 65536 independent tiles PER GPU with all operands unique ("mode S" of SURVEY.md 8d). One step is
one pass over the batch = ONE launch of the tcgen05 tile kernel through the C ABI
(libxsmm_b200_gemm_batch_strided). Inputs (8.6 GB) are far larger than L2, so no flush is needed.

  value  whole-job GFLOP/s, operands resident in HBM, CUDA events on the launching stream, max over ranks
  e2e    same metric through the same C-ABI call with HOST (pinned) buffers: H2D + kernel + D2H per step
  roofline      achieved algorithmic GB/s of the dominant kernel vs the measured HBM peak
  cpu_baseline  the reference's own JIT kernel (oracle/_ref, all host cores) on a bounded sample
  also          fsspmdm (config 3) and BCSC (config 4) with their own roofline numbers

--impl reference runs the unmodified reference (AMX/AVX-512 JIT through its public dispatch API, OpenMP over
the batch) on a bounded sample of the same workload on rank 0 only.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

M = N = K = 64
BR = 8
BATCH = 65536
BF16, F32 = 2, 1
FLAG_BETA_0 = 4
WORKLOAD = ("configs[1]: batched BRGEMM bf16->f32 m=n=k=64 brcount=8 batch=65536 per GPU, stride-BR, beta=0, "
            "mode S (all operands unique)")
# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch from the committed `ncu --set full` captures (bytes)
NCU_TRAFFIC = {"gemm_tc_kernel<64>": (9652.8e6, "profiles/r01_ncu_gemm_tc.txt"),
               "sreg_kernel<float>": (622.2e6, "profiles/r01_ncu_sreg.txt"),
               "bcsc_tc_kernel<32>": (None, "profiles/r01_ncu_bcsc_tc.txt")}


def traffic(kernel):
    t = NCU_TRAFFIC.get(kernel, (None, None))
    return t[0]


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "src": "fallback"}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            j = json.load(f)
        p = {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops": float(j["bf16_tflops"]), "src": "measured"}
    except Exception:
        pass
    return p


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.rows, self.stop_flag, self.index = [], False, index
        self.thread = threading.Thread(target=self._run, daemon=True)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def __enter__(self):
        self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop_flag = True
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def fill_tenths(t, torch, chunk=1 << 26):
    """multiples of 0.1 in [-0.5, 0.5] like the reference drivers (spmm_kernel.c:498-527), seed 555"""
    g = torch.Generator(device=t.device); g.manual_seed(555)
    flat = t.view(-1)
    for s in range(0, flat.numel(), chunk):
        e = min(flat.numel(), s + chunk)
        flat[s:e] = (torch.randint(-5, 6, (e - s,), device=t.device, generator=g, dtype=torch.int8).to(torch.float32) / 10).to(t.dtype)


def time_steps(torch, fn, steps, warmup, dist=None):
    """W warm-up + K timed steps between barrier+synchronize; CUDA events on the launching stream; max over ranks"""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    ev[0].record()
    for i in range(steps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
    total_ms = ev[0].elapsed_time(ev[steps])
    if dist is not None:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    return total_ms, per


# ------------------------------------------------------------------------------------------------ BRGEMM
def brgemm_setup(X, torch, batch, out_f32=True):
    tc = F32 if out_f32 else BF16
    shape = X.libxsmm_create_gemm_shape(M, N, K, M, K, M, BF16, BF16, tc, F32)
    cfg = X.libxsmm_create_gemm_batch_reduce_config(X.GEMM_BATCH_REDUCE_STRIDE, M * K * 2, K * N * 2, 0)
    kernel = X.libxsmm_dispatch_brgemm(shape, FLAG_BETA_0, 0, cfg)
    assert kernel, "dispatch failed"
    assert X.libxsmm_b200_kernel_backend(kernel) == X.BACKEND_TCGEN05, "tcgen05 kernel not selected"
    a = torch.empty(batch * BR * M * K, dtype=torch.bfloat16, device="cuda")
    b = torch.empty(batch * BR * K * N, dtype=torch.bfloat16, device="cuda")
    c = torch.empty(batch * M * N, dtype=torch.float32 if out_f32 else torch.bfloat16, device="cuda")
    fill_tenths(a, torch); fill_tenths(b, torch)
    strides = (BR * M * K * 2, BR * K * N * 2, M * N * (4 if out_f32 else 2))
    return kernel, a, b, c, strides


def brgemm_check(X, torch, a, b, c, strides, ntiles=4):
    """a few tiles against the CPU oracle (test infrastructure used as the checker only)"""
    import numpy as np
    from oracle_ffi import oracle, run_gemm
    import gen
    for t in (0, 1, 77, BATCH // 3)[:ntiles]:
        if (t + 1) * M * N > c.numel():
            continue
        ah = a[t * BR * M * K:(t + 1) * BR * M * K].view(torch.int16).cpu().numpy().view(np.uint16)
        bh = b[t * BR * K * N:(t + 1) * BR * K * N].view(torch.int16).cpu().numpy().view(np.uint16)
        want = np.zeros(M * N, dtype=np.float32)
        assert run_gemm(oracle, (M, N, K, M, K, M), (BF16, BF16, F32, F32), FLAG_BETA_0, 3, M * K * 2, K * N * 2, BR, ah, bh, want) == 0
        got = c[t * M * N:(t + 1) * M * N].float().cpu().numpy()
        err = gen.normf_rel(want, got)
        assert err < 1.2e-5, "bench output differs from the oracle (tile %d, err %g)" % (t, err)


def run_ours(args):
    import torch
    import libxsmm_b200 as X
    dist = None
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
    X.libxsmm_b200_set_device(local)
    X.libxsmm_b200_set_stream(torch.cuda.current_stream().cuda_stream)
    X.libxsmm_b200_set_blocking(0)
    pk = peaks()
    out = {}
    if args.workload == "brgemm":
        kernel, a, b, c, (sa, sb, sc) = brgemm_setup(X, torch, BATCH)

        def step():
            rc = X.libxsmm_b200_gemm_batch_strided(kernel, a.data_ptr(), b.data_ptr(), c.data_ptr(), sa, sb, sc, BR, BATCH)
            assert rc == 0, X.libxsmm_b200_last_error_string()
        launches0 = X.libxsmm_b200_launch_count()
        step(); X.check(); brgemm_check(X, torch, a, b, c, (sa, sb, sc))
        launches0 = X.libxsmm_b200_launch_count()
        with ClockSampler(local) as clocks:
            total_ms, per = time_steps(torch, step, args.steps, args.warmup, dist)
        launches = X.libxsmm_b200_launch_count() - launches0 - args.warmup
        X.check()
        from libxsmm_b200.shard import weak_batch
        per_gpu, job_tiles = weak_batch(BATCH, world)      # batch is the only shard axis: every rank owns BATCH tiles, no collective
        flops = 2.0 * M * N * K * BR * per_gpu
        bytes_alg = float(BATCH) * (sa + sb + sc)
        ms = total_ms / args.steps
        kern_ms = sorted(per)[len(per) // 2]
        value = (2.0 * M * N * K * BR * job_tiles) / (ms * 1e-3) / 1e9     # whole job; ms is the max over ranks
        ach = bytes_alg / (kern_ms * 1e-3) / 1e9
        out = {"metric": "batched BRGEMM GFLOP/s (bf16, m=n=k=64, br=8, batch=65536/GPU, unique operands)", "value": value, "unit": "GFLOP/s",
               "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
               "config": {"workload": WORKLOAD,
                          "batch_per_gpu": BATCH, "parallelism": "batch sharded, no collective", "l2_policy": "inputs (8.6 GB) larger than L2, no flush"},
               "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic("gemm_tc_kernel<64>"),
                            "traffic_src": NCU_TRAFFIC["gemm_tc_kernel<64>"][1], "algorithmic_bytes": bytes_alg,
                            "peak_src": pk["src"], "kernel": "gemm_tc_kernel<64>", "kernel_ms": kern_ms,
                            "tensor_frac_of_measured_bf16_peak": (flops / (kern_ms * 1e-3) / 1e12) / pk["bf16_tflops"]},
               "gpu_launches": int(launches), "clocks": clocks.summary()}
        if rank == 0 and not args.no_e2e:
            out["e2e"] = brgemm_e2e(X, torch, kernel, a, b, sa, sb, sc, flops, c_dev=c)
        if rank == 0 and not args.no_also:
            out["also"] = {}
            for name, fn in (("fsspmdm", also_fsspmdm), ("bcsc", also_bcsc)):
                try:
                    out["also"][name] = fn(X, torch, pk, args)
                except Exception as e:  # secondary numbers must not take the headline down
                    out["also"][name] = {"error": repr(e)[:200]}
        if rank == 0 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline_brgemm()
    elif args.workload == "fsspmdm":
        out = also_fsspmdm(X, torch, pk, args, full=True)
    elif args.workload == "bcsc":
        out = also_bcsc(X, torch, pk, args, full=True)
    elif args.workload == "sweep":
        out = sweep(X, torch, pk, args)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out))


def brgemm_e2e(X, torch, kernel, a, b, sa, sb, sc, flops, steps=3, c_dev=None):
    """same call, HOST pinned buffers: the library moves A and B to the device, runs the kernel and brings C back inside the
    step. Two transports are timed: the chunked three-stream copy pipeline (default) and in-place access to the pinned
    buffers from the kernel (LIBXSMM_B200_ZEROCOPY=1); the better one is the e2e value, both are reported."""
    nb_a, nb_b, nb_c = BATCH * sa, BATCH * sb, BATCH * sc
    ha = torch.empty(nb_a // 2, dtype=torch.bfloat16, pin_memory=True); hb = torch.empty(nb_b // 2, dtype=torch.bfloat16, pin_memory=True)
    hc = torch.empty(nb_c // 4, dtype=torch.float32, pin_memory=True)
    ha.copy_(a); hb.copy_(b)          # the synthetic operands, now resident on the host
    X.libxsmm_b200_set_blocking(1)
    res = {}
    for mode, env in (("copy_pipeline", None), ("zero_copy", "1")):
        if env is None:
            os.environ.pop("LIBXSMM_B200_ZEROCOPY", None)
        else:
            os.environ["LIBXSMM_B200_ZEROCOPY"] = env
        best = None
        for i in range(steps + 1):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            rc = X.libxsmm_b200_gemm_batch_strided(kernel, ha.data_ptr(), hb.data_ptr(), hc.data_ptr(), sa, sb, sc, BR, BATCH)
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            assert rc == 0, X.libxsmm_b200_last_error_string()
            if i > 0:
                best = dt if best is None else min(best, dt)
        res[mode] = best
        if c_dev is not None:      # the host result of this transport equals the device-resident run (same kernel, same inputs)
            n = 1 << 20
            assert torch.equal(hc[:n], c_dev[:n].cpu()) and torch.equal(hc[-n:], c_dev[-n:].cpu()), "e2e result differs (%s)" % mode
    os.environ.pop("LIBXSMM_B200_ZEROCOPY", None)
    X.libxsmm_b200_set_blocking(0)
    mode = min(res, key=res.get)
    best = res[mode]
    return {"value": flops / best / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(nb_a + nb_b), "d2h_bytes_per_step": int(nb_c),
            "ms_per_step": best * 1e3, "transport": mode, "ms_by_transport": {k: v * 1e3 for k, v in res.items()},
            "note": "pinned host A,B -> libxsmm_b200_gemm_batch_strided -> pinned host C; wall clock around the blocking call, best of %d" % steps}


def also_fsspmdm(X, torch, pk, args, full=False):
    import numpy as np
    Mf, Kf, Nf = 32, 128, 1000000
    rng = np.random.default_rng(555)
    a = ((rng.integers(-5, 6, size=Mf * Kf) / 10.0) * (rng.random(Mf * Kf) < 0.15)).astype(np.float32)
    nnz = int(np.count_nonzero(a))
    one = np.array([1.0], dtype=np.float32); zero = np.array([0.0], dtype=np.float32)
    h = X.libxsmm_fsspmdm_create(F32, Mf, Nf, Kf, Kf, Nf, Nf, one.ctypes.data, zero.ctypes.data, a.ctypes.data, 0, None)
    assert h
    b = torch.randn(Kf * Nf, device="cuda"); c = torch.empty(Mf * Nf, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def step():
        X.libxsmm_fsspmdm_execute(h, b.data_ptr(), c.data_ptr())
    steps = max(5, args.steps)
    total_ms, per = time_steps(torch, step, steps, 3)
    X.check()
    ms = sorted(per)[len(per) // 2]
    bytes_alg = 4.0 * (Kf * Nf + Mf * Nf)
    ach = bytes_alg / (ms * 1e-3) / 1e9
    X.libxsmm_fsspmdm_destroy(h)
    cpu = None
    if full or not getattr(args, "no_cpu", False):
        cpu = cpu_baseline_fsspmdm(a, Mf, Kf, nnz)
    return {"cpu_baseline": cpu, "metric": "fsspmdm GFLOP/s (f32 M=32 K=128 N=1e6, 15% nnz)", "value": 2.0 * nnz * Nf / (ms * 1e-3) / 1e9, "unit": "GFLOP/s (sparse)",
            "dense_equiv_gflops": 2.0 * Mf * Kf * Nf / (ms * 1e-3) / 1e9, "ms_per_step": ms, "nnz": nnz,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic("sreg_kernel<float>"), "algorithmic_bytes": bytes_alg, "kernel": "sreg_kernel<float>"},
            "config": {"workload": "configs[2]: fsspmdm f32 M=32 K=128 N=1e6 15% nnz beta=0; B+C = 640 MB per step (> L2)"}}


def also_bcsc(X, torch, pk, args, full=False, mblocks=8192):
    import numpy as np
    Mb, Kb, Nb, bk, bn = 32, 512, 512, 32, 32
    rng = np.random.default_rng(555)
    nbr, nbc = Kb // bk, Nb // bn
    keep = np.zeros(nbr * nbc, dtype=bool); keep[rng.permutation(nbr * nbc)[:nbr * nbc // 2]] = True
    keep = keep.reshape(nbc, nbr)
    colptr = np.concatenate([[0], np.cumsum(keep.sum(1))]).astype(np.uint32); rowidx = np.nonzero(keep)[1].astype(np.uint32)
    nnzb = int(colptr[-1])
    shape = X.libxsmm_create_gemm_shape(mblocks, 0, Kb, Kb, 0, Nb, BF16, BF16, BF16, F32)
    kernel = X.libxsmm_create_packed_spgemm_bcsc(shape, FLAG_BETA_0 | X.GEMM_FLAG_VNNI_A, 0, X.SpgemmConfig(Mb, bk, bn))
    assert kernel
    a = torch.empty(mblocks * Kb * Mb, dtype=torch.bfloat16, device="cuda"); fill_tenths(a, torch)
    bv = torch.empty(nnzb * bk * bn, dtype=torch.bfloat16, device="cuda"); fill_tenths(bv, torch)
    c = torch.empty(mblocks * Nb * Mb, dtype=torch.bfloat16, device="cuda")
    d_cp = torch.from_numpy(colptr.view(np.int32).copy()).cuda(); d_ri = torch.from_numpy(rowidx.view(np.int32).copy()).cuda()
    p = X.GemmParam(); nb = C.c_ulonglong(nbc)
    p.a.primary, p.b.primary, p.b.secondary, p.b.tertiary, p.b.quaternary, p.c.primary = a.data_ptr(), bv.data_ptr(), d_cp.data_ptr(), d_ri.data_ptr(), C.addressof(nb), c.data_ptr()
    fn = X.GEMMFUNCTION(kernel)

    def step():
        fn(C.byref(p))
    steps = max(3, args.steps // 4)
    total_ms, per = time_steps(torch, step, steps, 2)
    X.check()
    ms = sorted(per)[len(per) // 2]
    bytes_alg = 2.0 * (mblocks * Kb * Mb + mblocks * Nb * Mb) + 2.0 * nnzb * bk * bn
    ach = bytes_alg / (ms * 1e-3) / 1e9
    X.libxsmm_release_kernel(kernel)
    cpu = None
    if mblocks == 8192 and (full or not getattr(args, "no_cpu", False)):
        cpu = cpu_baseline_bcsc(colptr, rowidx, nnzb, (Mb, Kb, Nb, bk, bn))
    return {"cpu_baseline": cpu, "metric": "BCSC spmm GFLOP/s dense-equivalent (bf16, M=32 N=K=512, 32x32 blocks, 50%, m_blocks=8192)",
            "value": 2.0 * Mb * mblocks * Nb * Kb / (ms * 1e-3) / 1e9, "unit": "GFLOP/s (dense-equivalent)",
            "effective_gflops": 2.0 * Mb * mblocks * nnzb * bk * bn / (ms * 1e-3) / 1e9, "ms_per_step": ms,
            "roofline": {"bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"], "traffic": traffic("bcsc_tc_kernel<32>"), "algorithmic_bytes": bytes_alg,
                         "kernel": "bcsc_tc_kernel<32> (+ bcsc_prep_kernel, bcsc_pack_b_kernel: the timed call is all three launches)"},
            "config": {"workload": "configs[3] on one GPU: BCSC bf16 M=32 N=K=512 bk=bn=32 50% m_blocks=8192; A+C = 537 MB per step (> L2)"}}


def sweep(X, torch, pk, args, batch=32768):
    """configs[4]: int8 x int8 -> int32 (U8 x I8, VNNI4 A) and F16 x F16 -> F32, m=n=k in {8..128}, br=1, unique operands"""
    I8, U8, I32, F16 = 12, 13, 8, 3
    pts = []
    for name, ta, tb, tcc, tcomp, flags, esz, csz in (("u8*i8->i32", U8, I8, I32, I32, FLAG_BETA_0 | X.GEMM_FLAG_VNNI_A, 1, 4),
                                                       ("f16*f16->f32", F16, F16, F32, F32, FLAG_BETA_0, 2, 4)):
        for m in (8, 16, 32, 64, 128):
            shape = X.libxsmm_create_gemm_shape(m, m, m, m, m, m, ta, tb, tcc, tcomp)
            kernel = X.libxsmm_dispatch_gemm(shape, flags, 0)
            if not kernel:
                pts.append({"type": name, "m": m, "error": "dispatch returned NULL"}); continue
            a = torch.randint(0, 5, (batch * m * m * esz,), dtype=torch.uint8, device="cuda")
            b = torch.randint(0, 5, (batch * m * m * esz,), dtype=torch.uint8, device="cuda")
            if esz == 2:
                a = (torch.randint(-5, 6, (batch * m * m,), device="cuda").float() / 10).half(); b = a.roll(7)
            c = torch.empty(batch * m * m * csz, dtype=torch.uint8, device="cuda")
            sa = sb = m * m * esz; sc = m * m * csz

            def step():
                rc = X.libxsmm_b200_gemm_batch_strided(kernel, a.data_ptr(), b.data_ptr(), c.data_ptr(), sa, sb, sc, 1, batch)
                assert rc == 0, X.libxsmm_b200_last_error_string()
            total_ms, per = time_steps(torch, step, max(5, args.steps // 2), 3)
            X.check()
            ms = sorted(per)[len(per) // 2]
            bytes_alg = float(batch) * (sa + sb + sc)
            ach = bytes_alg / (ms * 1e-3) / 1e9
            pts.append({"type": name, "m": m, "gflops": 2.0 * m * m * m * batch / (ms * 1e-3) / 1e9, "ms": ms, "gbs": ach, "hbm_frac": ach / pk["hbm_gbs"],
                        "backend": int(X.libxsmm_b200_kernel_backend(kernel)), "l2_note": "operands %.0f MB%s" % (bytes_alg / 1e6, "" if bytes_alg > 2.5e8 else " (fits L2: not an HBM number)")})
    best = max((p for p in pts if "gflops" in p), key=lambda p: p["gflops"])
    return {"metric": "mixed-precision sweep GFLOP/s (configs[4], diagonal m=n=k)", "value": best["gflops"], "unit": "GFLOP/s", "n_gpus": 1, "steps": args.steps,
            "warmup": 3, "higher_is_better": True, "dtype": "u8/i8->i32, f16->f32", "data": "synthetic",
            "config": {"workload": "configs[4]: int8 and f16 GEMM m=n=k in {8,16,32,64,128}, batch=32768, br=1, beta=0, unique operands"},
            "points": pts, "roofline": {"bound": "hbm", "achieved": best["gbs"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": best["hbm_frac"], "traffic": None}}


# ------------------------------------------------------------------------------------------------ CPU side
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1; the CPU arm is meant to use every core this process may run on.
    Must run before the OpenMP runtime of oracle/_ref is loaded."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.setdefault("OMP_PROC_BIND", "false")
    return n


def cpu_baseline_brgemm(sample_tiles=4096, budget_s=12.0):
    """the reference's own JIT BRGEMM kernel over a bounded sample of the same batch, all host cores (OpenMP)"""
    import numpy as np
    use_all_host_threads()
    from oracle_ffi import ref_lib, iarr
    if ref_lib is None:
        return {"value": None, "unit": "GFLOP/s", "cores": 0, "kind": "reference", "sample": "oracle/_ref/libxsmm_ref.so missing"}
    rng = np.random.default_rng(555)
    na = sample_tiles * BR * M * K
    a = (rng.integers(-5, 6, size=na).astype(np.float32) / 10).view(np.uint32)
    a = ((a + 0x7FFF + ((a >> 16) & 1)) >> 16).astype(np.uint16)
    b = np.roll(a, 12345).copy()
    c = np.zeros(sample_tiles * M * N, dtype=np.float32)
    dims, types = iarr(M, N, K, M, K, M), iarr(BF16, BF16, F32, F32)
    is_ref = C.c_int(0)
    flags = FLAG_BETA_0 | 256   # VNNI_A: required by the reference's x86 bf16 JIT (AMX/AVX-512 BF16)
    args = (dims, types, flags, 3, M * K * 2, K * N * 2, BR, a.ctypes.data, b.ctypes.data, c.ctypes.data, BR * M * K * 2, BR * K * N * 2, M * N * 4, sample_tiles)
    t1 = ref_lib.ref_bench_gemm_batch(*args, 1, C.byref(is_ref))
    if t1 < 0:
        return {"value": None, "unit": "GFLOP/s", "cores": int(ref_lib.ref_max_threads()), "kind": "reference", "sample": "JIT dispatch returned NULL on this host"}
    reps = max(1, min(200, int(budget_s / max(t1, 1e-4))))
    t = ref_lib.ref_bench_gemm_batch(*args, reps, C.byref(is_ref))
    fl = 2.0 * M * N * K * BR * sample_tiles * reps
    return {"value": fl / t / 1e9, "unit": "GFLOP/s", "cores": int(ref_lib.ref_max_threads()), "kind": "reference",
            "sample": "%d tiles (%.0f MB of A+B, streamed) x %d passes, LIBXSMM JIT target %s%s, VNNI_A layout" % (
                sample_tiles, (2.0 * na * 2) / 1e6, reps, ref_lib.ref_target_arch().decode(), " [C reference kernel!]" if is_ref.value else "")}


def cpu_baseline_fsspmdm(a_dense, Mf, Kf, nnz, n_sample=200000, budget_s=4.0):
    """the reference's fsspmdm JIT on a bounded N-slice of the same operator, all host cores (slices of N per thread)"""
    import numpy as np
    use_all_host_threads()
    from oracle_ffi import ref_lib
    if ref_lib is None:
        return {"value": None, "unit": "GFLOP/s (sparse)", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    rng = np.random.default_rng(7)
    b = rng.standard_normal(Kf * n_sample).astype(np.float32); c = np.zeros(Mf * n_sample, dtype=np.float32)
    one = np.array([1.0], dtype=np.float32); zero = np.array([0.0], dtype=np.float32)
    args = (F32, Mf, n_sample, Kf, Kf, one.ctypes.data, zero.ctypes.data, a_dense.ctypes.data, b.ctypes.data, c.ctypes.data)
    t1 = ref_lib.ref_bench_fsspmdm(*args, 1)
    if t1 < 0:
        return {"value": None, "unit": "GFLOP/s (sparse)", "cores": int(ref_lib.ref_max_threads()), "kind": "reference", "sample": "create returned NULL"}
    reps = max(1, min(500, int(budget_s / max(t1, 1e-4))))
    t = ref_lib.ref_bench_fsspmdm(*args, reps)
    return {"value": 2.0 * nnz * n_sample * reps / t / 1e9, "unit": "GFLOP/s (sparse)", "gbs": 4.0 * (Kf + Mf) * n_sample * reps / t / 1e9,
            "cores": int(ref_lib.ref_max_threads()), "kind": "reference", "sample": "N=%d columns (%.0f MB of B+C) x %d passes, LIBXSMM JIT" % (n_sample, 4.0 * (Kf + Mf) * n_sample / 1e6, reps)}


def cpu_baseline_bcsc(colptr, rowidx, nnzb, geo, mblocks=1024, budget_s=4.0):
    """the reference's BCSC JIT (AMX on SPR) on a bounded number of m_blocks, contiguous ranges per thread"""
    import numpy as np
    use_all_host_threads()
    from oracle_ffi import ref_lib, iarr
    Mb, Kb, Nb, bk, bn = geo
    if ref_lib is None:
        return {"value": None, "unit": "GFLOP/s (dense-equivalent)", "cores": 0, "kind": "reference", "sample": "oracle/_ref missing"}
    rng = np.random.default_rng(8)

    def bf16(n):
        x = (rng.integers(-5, 6, size=n).astype(np.float32) / 10).view(np.uint32)
        return ((x + 0x7FFF + ((x >> 16) & 1)) >> 16).astype(np.uint16)
    a = bf16(mblocks * Kb * Mb); bv = bf16(nnzb * bk * bn); c = np.zeros(mblocks * Nb * Mb, dtype=np.uint16)
    cp = colptr.copy(); ri = rowidx.copy()
    args = (iarr(BF16, BF16, F32, BF16), iarr(mblocks, Mb, Kb, Nb, bk, bn), FLAG_BETA_0 | 256, a.ctypes.data, bv.ctypes.data, cp.ctypes.data, ri.ctypes.data, c.ctypes.data)
    t1 = ref_lib.ref_bench_bcsc(*args, 1)
    if t1 < 0:
        return {"value": None, "unit": "GFLOP/s (dense-equivalent)", "cores": int(ref_lib.ref_max_threads()), "kind": "reference", "sample": "JIT returned NULL on this host"}
    reps = max(1, min(500, int(budget_s / max(t1, 1e-4))))
    t = ref_lib.ref_bench_bcsc(*args, reps)
    return {"value": 2.0 * Mb * mblocks * Nb * Kb * reps / t / 1e9, "unit": "GFLOP/s (dense-equivalent)", "cores": int(ref_lib.ref_max_threads()), "kind": "reference",
            "sample": "%d m_blocks (%.0f MB of A+C) x %d passes, LIBXSMM JIT target %s" % (mblocks, 4.0 * mblocks * Mb * (Kb + Nb) / 2 / 1e6, reps, ref_lib.ref_target_arch().decode())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    vals = []
    t0 = time.perf_counter()
    for _ in range(args.warmup + args.steps):
        vals.append(cpu_baseline_brgemm(sample_tiles=2048, budget_s=1.5))
    dt = time.perf_counter() - t0
    good = [v["value"] for v in vals[args.warmup:] if v["value"]]
    if not good:
        print(json.dumps({"impl": "reference", "unavailable": vals[-1]["sample"]}))
        return
    v = sorted(good)[len(good) // 2]
    base = vals[-1]
    print(json.dumps({"impl": "reference", "metric": "batched BRGEMM GFLOP/s (bf16, m=n=k=64, br=8, batch=65536/GPU, unique operands)",
                      "value": v, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                      "ms_per_step": dt / (args.warmup + args.steps) * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "bf16", "data": "synthetic",
                      "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH,
                                 "reference_arm": "host CPU: each step a bounded sample (2048 tiles) of the batch, LIBXSMM JIT through libxsmm_dispatch_brgemm, OpenMP over tiles"},
                      "cpu_baseline": {"value": v, "unit": "GFLOP/s", "cores": base["cores"], "kind": "reference", "sample": base["sample"]},
                      "e2e": {"value": v, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="brgemm", choices=["brgemm", "fsspmdm", "bcsc", "sweep"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
