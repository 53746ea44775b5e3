"""GPU parity of the matrix-eltwise (TPP) kernels against the reference's portable C kernels
(libxsmm_reference_elementwise, via oracle/_ref) on seeded inputs; sizes like the reference's own
eltwise suite (m, n in 1..100, eqld and gtld). Data movement, masks, integer outputs and IEEE-exact
ops are compared bit for bit; transcendental ops with the bounds of samples/eltwise/eltwise_unary_simple.c:569-591."""
import ctypes as C

import numpy as np
import pytest
import torch

import gen
import libxsmm_b200 as X
from gpu_util import dev, host
from oracle_ffi import iarr, oracle, ref

pytestmark = [pytest.mark.gpu]

UNS = gen.F64 + 26   # LIBXSMM_DATATYPE_UNSUPPORTED
EXACT_UNARY = ["IDENTITY", "XOR", "X2", "SQRT", "NEGATE", "INC", "RECIPROCAL", "RECIPROCAL_SQRT"]
APPROX_UNARY = ["TANH", "TANH_INV", "SIGMOID", "SIGMOID_INV", "GELU", "GELU_INV", "EXP"]


def _rand(rng, n, t, positive=False):
    x = rng.standard_normal(n).astype(np.float32)
    if positive:
        x = np.abs(x) + 0.1
    if t == gen.F32:
        return x
    if t == gen.F64:
        return x.astype(np.float64)
    if t == gen.BF16:
        return gen.f32_to_bf16_bits(x)
    return x.astype(np.float16).view(np.uint16)


def _ref_call(desc, param):
    """expected result: the reference itself where oracle/_ref travelled with the snapshot, else the pinned restatement
    (oracle/oracle_meltw.c; it answers 2 for the operations it does not restate)"""
    if ref is not None:
        assert ref["meltw"](iarr(*desc), C.addressof(param), 0) == 0
        return
    rc = oracle["meltw"](iarr(*desc), C.addressof(param), 0)
    if rc == 2:
        pytest.skip("operation not restated in oracle/oracle_meltw.c and oracle/_ref is absent")
    assert rc == 0


def _desc(op_class, op, flags, m, n, ldi, ldi2, ldi3, ldo, t0, t1, t2, to, tcomp):
    return (op_class, op, flags, m, n, ldi, ldi2, ldi3, ldo, t0, t1, t2, to, tcomp)


def _cmp(got, want, t, exact, tol):
    if exact:
        assert np.array_equal(got.view(np.uint8), want.view(np.uint8))
    else:
        g, w = gen.to_f64(got, t), gen.to_f64(want, t)
        ok = np.isfinite(w)
        assert np.allclose(g[ok], w[ok], rtol=tol, atol=tol)


@pytest.mark.parametrize("tin,tout", [(gen.F32, gen.F32), (gen.BF16, gen.BF16), (gen.F16, gen.F32), (gen.F32, gen.BF16), (gen.F64, gen.F64)])
def test_unary_generic_ops(tin, tout):
    rng = np.random.default_rng(51)
    names = EXACT_UNARY + ([] if tin == gen.F64 else APPROX_UNARY)
    for name in names:
        op = getattr(X, "MELTW_TYPE_UNARY_" + name)
        for (m, n, pad, bc) in ((33, 17, 0, 0), (100, 3, 5, 0), (1, 64, 2, 0), (40, 9, 0, X.MELTW_FLAG_UNARY_BCAST_ROW),
                                (40, 9, 0, X.MELTW_FLAG_UNARY_BCAST_COL), (7, 7, 1, X.MELTW_FLAG_UNARY_BCAST_SCALAR)):
            ldi, ldo = m + pad, m + 2 * pad
            x = _rand(rng, ldi * n, tin, positive=name in ("SQRT", "RECIPROCAL", "RECIPROCAL_SQRT"))
            y0 = _rand(rng, ldo * n, tout)
            tcomp = gen.F64 if tin == gen.F64 else gen.F32
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, tin, tout, tcomp), bc)
            assert k, (name, tin, tout)
            d_x, d_y = dev(x), dev(y0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_y.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = y0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
            _ref_call(_desc(1, op, bc, m, n, ldi, 0, 0, ldo, tin, UNS, UNS, tout, tcomp), q)
            tol = 7e-4 if tout == gen.F32 else 7e-3
            _cmp(host(d_y, gen.NP_OF[tout]), want, tout, name in EXACT_UNARY, tol)


@pytest.mark.parametrize("t", [gen.F32, gen.BF16])
def test_relu_family_with_bitmask(t):
    rng = np.random.default_rng(52)
    for fwd, inv in (("RELU", "RELU_INV"), ("LEAKY_RELU", "LEAKY_RELU_INV"), ("ELU", "ELU_INV")):
        for (m, n, pad) in ((35, 11, 0), (64, 5, 3), (9, 40, 7)):
            for bitm in ((1, 0) if fwd != "ELU" else (0,)):
                ld = m + pad
                flags = X.MELTW_FLAG_UNARY_BITMASK_2BYTEMULT if bitm else 0
                x = _rand(rng, ld * n, t); y0 = _rand(rng, ld * n, t)
                alpha = C.c_float(0.3)
                mask_ld = (ld + 15) // 16 * 16
                mask0 = rng.integers(0, 256, size=mask_ld // 8 * n, dtype=np.uint8)
                op = getattr(X, "MELTW_TYPE_UNARY_" + fwd)
                k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, t, t, gen.F32), flags)
                assert k
                d_x, d_y, d_m = dev(x), dev(y0), dev(mask0)
                p = X.MeltwUnaryParam(); p.op.primary = C.addressof(alpha)
                p.inp.primary, p.out.primary, p.out.secondary = d_x.data_ptr(), d_y.data_ptr(), d_m.data_ptr()
                X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
                want = y0.copy(); wmask = mask0.copy()
                q = X.MeltwUnaryParam(); q.op.primary = C.addressof(alpha)
                q.inp.primary, q.out.primary, q.out.secondary = x.ctypes.data, want.ctypes.data, wmask.ctypes.data
                _ref_call(_desc(1, op, flags, m, n, ld, 0, 0, ld, t, UNS, UNS, t, gen.F32), q)
                _cmp(host(d_y, gen.NP_OF[t]), want, t, fwd != "ELU", 7e-3 if t == gen.BF16 else 7e-4)
                if bitm:
                    assert np.array_equal(host(d_m, np.uint8), wmask), (fwd, m, n)
                # backward: needs the mask (relu/leaky) or the forward output (elu)
                if fwd == "ELU" or bitm:
                    g = _rand(rng, ld * n, t); o0 = _rand(rng, ld * n, t)
                    opi = getattr(X, "MELTW_TYPE_UNARY_" + inv)
                    ki = X.libxsmm_dispatch_meltw_unary(opi, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, t, t, gen.F32), flags)
                    assert ki
                    aux_h = want if fwd == "ELU" else wmask
                    d_g, d_o, d_aux = dev(g), dev(o0), dev(aux_h)
                    p = X.MeltwUnaryParam(); p.op.primary = C.addressof(alpha)
                    p.inp.primary, p.inp.secondary, p.out.primary = d_g.data_ptr(), d_aux.data_ptr(), d_o.data_ptr()
                    X.MELTW_UNARY_FN(ki)(C.byref(p)); X.check()
                    wo = o0.copy(); q = X.MeltwUnaryParam(); q.op.primary = C.addressof(alpha)
                    q.inp.primary, q.inp.secondary, q.out.primary = g.ctypes.data, aux_h.ctypes.data, wo.ctypes.data
                    _ref_call(_desc(1, opi, flags, m, n, ld, 0, 0, ld, t, UNS, UNS, t, gen.F32), q)
                    _cmp(host(d_o, gen.NP_OF[t]), wo, t, True, 0)


@pytest.mark.parametrize("t", [gen.F32, gen.BF16, gen.F64])
def test_binary_and_ternary(t):
    rng = np.random.default_rng(53)
    tcomp = gen.F64 if t == gen.F64 else gen.F32
    for name in ("ADD", "MUL", "SUB", "DIV", "MULADD", "MAX", "MIN"):
        op = getattr(X, "MELTW_TYPE_BINARY_" + name)
        for (m, n, pad, fl) in ((33, 17, 0, 0), (50, 4, 3, X.MELTW_FLAG_BINARY_BCAST_COL_IN_0), (20, 20, 0, X.MELTW_FLAG_BINARY_BCAST_ROW_IN_1),
                                (8, 3, 1, X.MELTW_FLAG_BINARY_BCAST_SCALAR_IN_1)):
            ld = m + pad
            a, b, o0 = _rand(rng, ld * n, t), _rand(rng, ld * n, t, positive=(name == "DIV")), _rand(rng, ld * n, t)
            k = X.libxsmm_dispatch_meltw_binary(op, X.libxsmm_create_meltw_binary_shape(m, n, ld, ld, ld, t, t, t, tcomp), fl)
            assert k, name
            d_a, d_b, d_o = dev(a), dev(b), dev(o0)
            p = X.MeltwBinaryParam(); p.in0.primary, p.in1.primary, p.out.primary = d_a.data_ptr(), d_b.data_ptr(), d_o.data_ptr()
            X.MELTW_BINARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwBinaryParam(); q.in0.primary, q.in1.primary, q.out.primary = a.ctypes.data, b.ctypes.data, want.ctypes.data
            _ref_call(_desc(2, op, fl, m, n, ld, ld, 0, ld, t, t, UNS, t, tcomp), q)
            _cmp(host(d_o, gen.NP_OF[t]), want, t, True, 0)
    if t == gen.F64:
        return
    # compare -> bitmask, then select by that bitmask
    for name in ("GT", "GE", "LT", "LE", "EQ", "NE"):
        op = getattr(X, "MELTW_TYPE_BINARY_CMP_OP_" + name)
        m, n, ld = 37, 9, 40
        a, b = _rand(rng, ld * n, t), _rand(rng, ld * n, t)
        b[::5] = a[::5]
        mask_ld = (ld + 15) // 16 * 16
        mask0 = rng.integers(0, 256, size=mask_ld // 8 * n, dtype=np.uint8)
        k = X.libxsmm_dispatch_meltw_binary(op, X.libxsmm_create_meltw_binary_shape(m, n, ld, ld, ld, t, t, t, gen.F32), X.MELTW_FLAG_BINARY_BITMASK_2BYTEMULT)
        assert k
        d_a, d_b, d_m = dev(a), dev(b), dev(mask0)
        p = X.MeltwBinaryParam(); p.in0.primary, p.in1.primary, p.out.primary = d_a.data_ptr(), d_b.data_ptr(), d_m.data_ptr()
        X.MELTW_BINARY_FN(k)(C.byref(p)); X.check()
        wmask = mask0.copy(); q = X.MeltwBinaryParam(); q.in0.primary, q.in1.primary, q.out.primary = a.ctypes.data, b.ctypes.data, wmask.ctypes.data
        _ref_call(_desc(2, op, X.MELTW_FLAG_BINARY_BITMASK_2BYTEMULT, m, n, ld, ld, 0, ld, t, t, UNS, t, gen.F32), q)
        assert np.array_equal(host(d_m, np.uint8), wmask), name
        ks = X.libxsmm_dispatch_meltw_ternary(X.MELTW_TYPE_TERNARY_SELECT, X.libxsmm_create_meltw_ternary_shape(m, n, ld, ld, ld, ld, t, t, gen.F32 + 99 if False else t, t, gen.F32),
                                              X.MELTW_FLAG_TERNARY_BITMASK_2BYTEMULT)
        assert ks
        o0 = _rand(rng, ld * n, t); d_o = dev(o0)
        p = X.MeltwTernaryParam(); p.in0.primary, p.in1.primary, p.in2.primary, p.out.primary = d_a.data_ptr(), d_b.data_ptr(), d_m.data_ptr(), d_o.data_ptr()
        X.MELTW_TERNARY_FN(ks)(C.byref(p)); X.check()
        want = o0.copy(); q = X.MeltwTernaryParam()
        q.in0.primary, q.in1.primary, q.in2.primary, q.out.primary = a.ctypes.data, b.ctypes.data, wmask.ctypes.data, want.ctypes.data
        _ref_call(_desc(3, X.MELTW_TYPE_TERNARY_SELECT, X.MELTW_FLAG_TERNARY_BITMASK_2BYTEMULT, m, n, ld, ld, ld, ld, t, t, t, t, gen.F32), q)
        _cmp(host(d_o, gen.NP_OF[t]), want, t, True, 0)
    for op in (X.MELTW_TYPE_TERNARY_MULADD, X.MELTW_TYPE_TERNARY_NMULADD):
        m, n, ld = 21, 13, 24
        a, b, c_, o0 = (_rand(rng, ld * n, t) for _ in range(4))
        k = X.libxsmm_dispatch_meltw_ternary(op, X.libxsmm_create_meltw_ternary_shape(m, n, ld, ld, ld, ld, t, t, t, t, gen.F32), 0)
        assert k
        d_a, d_b, d_c, d_o = dev(a), dev(b), dev(c_), dev(o0)
        p = X.MeltwTernaryParam(); p.in0.primary, p.in1.primary, p.in2.primary, p.out.primary = d_a.data_ptr(), d_b.data_ptr(), d_c.data_ptr(), d_o.data_ptr()
        X.MELTW_TERNARY_FN(k)(C.byref(p)); X.check()
        want = o0.copy(); q = X.MeltwTernaryParam()
        q.in0.primary, q.in1.primary, q.in2.primary, q.out.primary = a.ctypes.data, b.ctypes.data, c_.ctypes.data, want.ctypes.data
        _ref_call(_desc(3, op, 0, m, n, ld, ld, ld, ld, t, t, t, t, gen.F32), q)
        _cmp(host(d_o, gen.NP_OF[t]), want, t, True, 0)


@pytest.mark.parametrize("t", [gen.F32, gen.BF16, gen.F64])
def test_reductions(t):
    rng = np.random.default_rng(54)
    tcomp = gen.F64 if t == gen.F64 else gen.F32
    for name in ("REDUCE_X_OP_ADD", "REDUCE_X2_OP_ADD", "REDUCE_X_X2_OP_ADD", "REDUCE_X_OP_MAX", "REDUCE_X_OP_MIN", "REDUCE_X_OP_ABSMAX"):
        op = getattr(X, "MELTW_TYPE_UNARY_" + name)
        for rows in (1, 0):
            for (m, n, pad) in ((33, 17, 0), (100, 40, 4), (5, 77, 1)):
                ldi = m + pad
                nres = n if rows else m
                ldo = nres
                flags = X.MELTW_FLAG_UNARY_REDUCE_ROWS if rows else X.MELTW_FLAG_UNARY_REDUCE_COLS
                x = _rand(rng, ldi * n, t); o0 = _rand(rng, 2 * ldo, t)
                k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, tcomp), flags)
                assert k, name
                d_x, d_o = dev(x), dev(o0)
                p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
                X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
                want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
                _ref_call(_desc(1, op, flags, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, tcomp), q)
                got = host(d_o, gen.NP_OF[t])
                nvalid = nres * (2 if name == "REDUCE_X_X2_OP_ADD" else 1)
                exact = "ADD" not in name
                if name == "REDUCE_X2_OP_ADD" and t == gen.F64:
                    continue   # the reference's f64 x^2 path stores nothing (generator_mateltwise_reference_impl.c:1297)
                g, w = got[:nvalid], want[:nvalid]
                if name == "REDUCE_X_X2_OP_ADD" and t == gen.F64:
                    g, w = got[:nres], want[:nres]
                _cmp(g, w, t, exact, 1e-2 if t == gen.BF16 else 1e-4)


def test_transforms_and_gather_scatter_bit_exact():
    rng = np.random.default_rng(55)
    # --- transposes of every element size ---
    for t in (gen.F64, gen.F32, gen.BF16, gen.I8):
        for (m, n, pi, po) in ((33, 17, 0, 0), (64, 64, 2, 5), (1, 9, 0, 0)):
            ldi, ldo = m + pi, n + po
            x = rng.integers(0, 256, size=ldi * n * gen.TS[t], dtype=np.uint8); o0 = rng.integers(0, 256, size=ldo * m * gen.TS[t], dtype=np.uint8)
            op = X.MELTW_TYPE_UNARY_TRANSFORM_NORM_TO_NORMT
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, t), 0)
            assert k
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
            _ref_call(_desc(1, op, 0, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, t), q)
            assert np.array_equal(host(d_o, np.uint8), want), ("normt", t, m, n)
    # --- VNNI packers (even / multiple-of-4 n so that the reference does not read past the matrix) ---
    for name, t, v in (("NORM_TO_VNNI2", gen.BF16, 2), ("NORM_TO_VNNI4", gen.I8, 4), ("NORM_TO_VNNI4", gen.BF16, 4), ("NORM_TO_VNNI2T", gen.BF16, 2),
                       ("NORM_TO_VNNI4T", gen.BF16, 4), ("VNNI2_TO_VNNI2T", gen.BF16, 2), ("VNNI4_TO_VNNI4T", gen.I8, 4), ("VNNI4_TO_VNNI4T", gen.BF16, 4),
                       ("VNNI2T_TO_NORM", gen.BF16, 2), ("VNNI4T_TO_NORM", gen.BF16, 4), ("VNNI4_TO_NORM", gen.I8, 4)):
        op = getattr(X, "MELTW_TYPE_UNARY_TRANSFORM_" + name)
        for (m, n, pad) in ((32, 16, 0), (64, 8, 4), (8, 64, 0), (40, 12, 8)):
            ldi = m + pad
            ldo = (n if name in ("VNNI2_TO_VNNI2T", "VNNI4_TO_VNNI4T", "NORM_TO_VNNI2T", "NORM_TO_VNNI4T") else
                   (n if name in ("VNNI2T_TO_NORM", "VNNI4T_TO_NORM") else m)) + pad
            x = rng.integers(0, 256, size=(ldi + 8) * (n + 8) * 4 * gen.TS[t], dtype=np.uint8)
            o0 = rng.integers(0, 256, size=(ldo + 8) * (max(m, n) + 8) * 4 * gen.TS[t], dtype=np.uint8)
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, t), 0)
            assert k, name
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
            _ref_call(_desc(1, op, 0, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, t), q)
            assert np.array_equal(host(d_o, np.uint8), want), (name, t, m, n, pad)
    # --- gather / scatter ---
    for t in (gen.F32, gen.BF16, gen.I8):
        for mode, fl in (("cols", X.MELTW_FLAG_UNARY_GS_COLS), ("rows", X.MELTW_FLAG_UNARY_GS_ROWS), ("offs", X.MELTW_FLAG_UNARY_GS_OFFS)):
            for idx8 in (0, 1):
                m, n, big = 24, 10, 50
                it = np.uint64 if idx8 else np.uint32
                flags = fl | (X.MELTW_FLAG_UNARY_IDX_SIZE_8BYTES if idx8 else X.MELTW_FLAG_UNARY_IDX_SIZE_4BYTES)
                for op in (X.MELTW_TYPE_UNARY_GATHER, X.MELTW_TYPE_UNARY_SCATTER):
                    gather = op == X.MELTW_TYPE_UNARY_GATHER
                    if mode == "cols":
                        idx = rng.permutation(big)[:n].astype(it); ldi, ldo = m, m; in_n, out_n = (big, n) if gather else (n, big); in_m = out_m = m
                    elif mode == "rows":
                        idx = rng.permutation(big)[:m].astype(it); ldi, ldo = (big, m) if gather else (m, big); in_n = out_n = n
                    else:
                        idx = rng.permutation(big * big)[:m * n].astype(it); ldi, ldo = (big, m) if gather else (m, big); in_n, out_n = (big, n) if gather else (n, big)
                    x = rng.integers(0, 256, size=big * big * gen.TS[t], dtype=np.uint8); o0 = rng.integers(0, 256, size=big * big * gen.TS[t], dtype=np.uint8)
                    k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, t), flags)
                    assert k
                    d_x, d_o, d_i = dev(x), dev(o0), dev(idx)
                    p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
                    q = X.MeltwUnaryParam(); want = o0.copy(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
                    if gather:
                        p.inp.secondary, q.inp.secondary = d_i.data_ptr(), idx.ctypes.data
                    else:
                        p.out.secondary, q.out.secondary = d_i.data_ptr(), idx.ctypes.data
                    X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
                    _ref_call(_desc(1, op, flags, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, t), q)
                    assert np.array_equal(host(d_o, np.uint8), want), (t, mode, idx8, gather)


def test_quant_dequant_and_scalar_reductions():
    rng = np.random.default_rng(56)
    m, n, ld = 37, 11, 40
    scf = C.c_float(12.5)
    x = (rng.standard_normal(ld * n) * 8).astype(np.float32)
    for tout, npdt in ((gen.I8, np.int8), (gen.I16, np.int16), (gen.I32, np.int32)):
        for fl in (0, X.MELTW_FLAG_UNARY_SIGN_SAT_QUANT):
            if tout == gen.I32 and fl:
                continue
            o0 = rng.integers(-100, 100, size=ld * n).astype(npdt)
            k = X.libxsmm_dispatch_meltw_unary(X.MELTW_TYPE_UNARY_QUANT, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, gen.F32, tout, gen.F32), fl)
            assert k
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.inp.secondary, p.out.primary = d_x.data_ptr(), C.addressof(scf), d_o.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.inp.secondary, q.out.primary = x.ctypes.data, C.addressof(scf), want.ctypes.data
            _ref_call(_desc(1, X.MELTW_TYPE_UNARY_QUANT, fl, m, n, ld, 0, 0, ld, gen.F32, UNS, UNS, tout, gen.F32), q)
            assert np.array_equal(host(d_o, npdt), want), (tout, fl)
            # and back
            kd = X.libxsmm_dispatch_meltw_unary(X.MELTW_TYPE_UNARY_DEQUANT, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, tout, gen.F32, gen.F32), 0)
            assert kd
            f0 = rng.standard_normal(ld * n).astype(np.float32); d_f = dev(f0); d_q = dev(want)
            p = X.MeltwUnaryParam(); p.inp.primary, p.inp.secondary, p.out.primary = d_q.data_ptr(), C.addressof(scf), d_f.data_ptr()
            X.MELTW_UNARY_FN(kd)(C.byref(p)); X.check()
            wf = f0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.inp.secondary, q.out.primary = want.ctypes.data, C.addressof(scf), wf.ctypes.data
            _ref_call(_desc(1, X.MELTW_TYPE_UNARY_DEQUANT, 0, m, n, ld, 0, 0, ld, tout, UNS, UNS, gen.F32, gen.F32), q)
            assert np.array_equal(host(d_f, np.float32), wf)
    # reduce-to-scalar and dot product (order of summation differs: tolerance)
    a, b = rng.standard_normal(ld * n).astype(np.float32), rng.standard_normal(ld * n).astype(np.float32)
    k = X.libxsmm_dispatch_meltw_unary(X.MELTW_TYPE_UNARY_REDUCE_TO_SCALAR_OP_ADD, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, gen.F32, gen.F32, gen.F32), 0)
    d_a, d_o = dev(a), dev(np.zeros(4, dtype=np.float32))
    p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_a.data_ptr(), d_o.data_ptr()
    X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
    assert abs(float(host(d_o, np.float32)[0]) - float(a.reshape(n, ld)[:, :m].sum(dtype=np.float64))) < 1e-3
    k = X.libxsmm_dispatch_meltw_binary(X.MELTW_TYPE_BINARY_MUL_AND_REDUCE_TO_SCALAR_OP_ADD, X.libxsmm_create_meltw_binary_shape(m, n, ld, ld, ld, gen.F32, gen.F32, gen.F32, gen.F32), 0)
    d_b = dev(b)
    p = X.MeltwBinaryParam(); p.in0.primary, p.in1.primary, p.out.primary = d_a.data_ptr(), d_b.data_ptr(), d_o.data_ptr()
    X.MELTW_BINARY_FN(k)(C.byref(p)); X.check()
    want = float((a.reshape(n, ld)[:, :m].astype(np.float64) * b.reshape(n, ld)[:, :m]).sum())
    assert abs(float(host(d_o, np.float32)[0]) - want) < 1e-3


def test_vnni8_pad_and_vnni4_to_vnni2_transforms_bit_exact():
    rng = np.random.default_rng(58)
    cases_ = [("NORM_TO_VNNI8", gen.BF16, "m"), ("NORM_TO_VNNI8", gen.I8, "m"), ("NORM_TO_VNNI8_PAD", gen.BF16, "m"), ("NORM_TO_VNNI8T", gen.BF16, "n"),
              ("VNNI8_TO_VNNI8T", gen.BF16, "n"), ("VNNI8_TO_VNNI8T", gen.I8, "n"), ("VNNI8T_TO_NORM", gen.BF16, "n"), ("VNNI4_TO_VNNI2", gen.I8, "m"),
              ("PADM_MOD2", gen.BF16, "m"), ("PADN_MOD2", gen.BF16, "m"), ("PADNM_MOD2", gen.BF16, "m"),
              ("PADM_MOD4", gen.I8, "m"), ("PADN_MOD4", gen.I8, "m"), ("PADNM_MOD4", gen.I8, "m")]
    for name, t, ld_of in cases_:
        op = getattr(X, "MELTW_TYPE_UNARY_TRANSFORM_" + name)
        shapes = ((32, 16, 0), (64, 8, 8), (8, 64, 0), (40, 24, 8))
        if name.startswith("PAD"):
            shapes = shapes + ((33, 7, 3), (5, 9, 1))
        for (m, n, pad) in shapes:
            ldi, ldo = m + pad, (n if ld_of == "n" else m) + pad
            x = rng.integers(0, 256, size=(ldi + 8) * (n + 16) * 8 * gen.TS[t], dtype=np.uint8)
            o0 = rng.integers(0, 256, size=(ldo + 8) * (max(m, n) + 16) * 8 * gen.TS[t], dtype=np.uint8)
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, t), 0)
            assert k, name
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
            _ref_call(_desc(1, op, 0, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, t), q)
            assert np.array_equal(host(d_o, np.uint8), want), (name, t, m, n, pad)


@pytest.mark.parametrize("tin,tout", [(gen.F32, gen.F32), (gen.BF16, gen.BF16), (gen.F32, gen.BF16)])
def test_dropout_same_numbers_as_the_reference_generator(tin, tout):
    """forward: the 16-lane generator state is consumed column by column like the reference, so outputs, bitmask AND the
    advanced state must be bit-identical; backward: mask-driven"""
    rng = np.random.default_rng(59)
    for (m, n, pad) in ((33, 7, 0), (64, 5, 3), (100, 13, 4), (7, 9, 1)):
        for bitm in (0, X.MELTW_FLAG_UNARY_BITMASK_2BYTEMULT):
            ldi, ldo = m + pad, m + 2 * pad
            x = _rand(rng, ldi * n, tin); y0 = _rand(rng, ldo * n, tout)
            prob = C.c_float(0.3)
            mask0 = rng.integers(0, 256, size=((ldo + 15) // 16 * 16) // 8 * n + 8, dtype=np.uint8)
            st0 = rng.integers(1, 2**32 - 1, size=64, dtype=np.uint64).astype(np.uint32)
            op = X.MELTW_TYPE_UNARY_DROPOUT
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, tin, tout, gen.F32), bitm)
            assert k
            d_x, d_y, d_m, d_s = dev(x), dev(y0), dev(mask0), dev(st0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary, p.out.secondary = d_x.data_ptr(), d_y.data_ptr(), d_m.data_ptr()
            p.op.primary, p.op.secondary = C.addressof(prob), d_s.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want, wmask, wst = y0.copy(), mask0.copy(), st0.copy()
            q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary, q.out.secondary = x.ctypes.data, want.ctypes.data, wmask.ctypes.data
            q.op.primary, q.op.secondary = C.addressof(prob), wst.ctypes.data
            _ref_call(_desc(1, op, bitm, m, n, ldi, 0, 0, ldo, tin, UNS, UNS, tout, gen.F32), q)
            assert np.array_equal(host(d_y, want.dtype).view(np.uint8), want.view(np.uint8)), ("dropout", m, n, bitm)
            assert np.array_equal(host(d_s, np.uint32), wst), "generator state after the call"
            if bitm:
                assert np.array_equal(host(d_m, np.uint8), wmask), "dropout mask"
                # host-resident operands (staging path) must give the same bits and advance the caller's state in place
                hy, hm, hs = y0.copy(), mask0.copy(), st0.copy()
                r = X.MeltwUnaryParam(); r.inp.primary, r.out.primary, r.out.secondary = x.ctypes.data, hy.ctypes.data, hm.ctypes.data
                r.op.primary, r.op.secondary = C.addressof(prob), hs.ctypes.data
                X.MELTW_UNARY_FN(k)(C.byref(r)); X.check()
                assert np.array_equal(hy.view(np.uint8), want.view(np.uint8)) and np.array_equal(hm, wmask) and np.array_equal(hs, wst)
                opi = X.MELTW_TYPE_UNARY_DROPOUT_INV
                ki = X.libxsmm_dispatch_meltw_unary(opi, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, tin, tout, gen.F32), bitm)
                assert ki
                maskb = rng.integers(0, 256, size=((ldi + 15) // 16 * 16) // 8 * n + 8, dtype=np.uint8)
                d_mb, d_y2 = dev(maskb), dev(y0)
                pb = X.MeltwUnaryParam(); pb.inp.primary, pb.inp.secondary, pb.out.primary = d_x.data_ptr(), d_mb.data_ptr(), d_y2.data_ptr()
                pb.op.primary = C.addressof(prob)
                X.MELTW_UNARY_FN(ki)(C.byref(pb)); X.check()
                wb = y0.copy(); qb = X.MeltwUnaryParam(); qb.inp.primary, qb.inp.secondary, qb.out.primary = x.ctypes.data, maskb.ctypes.data, wb.ctypes.data
                qb.op.primary = C.addressof(prob)
                _ref_call(_desc(1, opi, bitm, m, n, ldi, 0, 0, ldo, tin, UNS, UNS, tout, gen.F32), qb)
                assert np.array_equal(host(d_y2, wb.dtype).view(np.uint8), wb.view(np.uint8)), ("dropout_inv", m, n)


def test_unzip_and_decompose_bit_exact():
    rng = np.random.default_rng(60)
    for (m, n, pad) in ((33, 7, 0), (64, 5, 3), (1, 9, 2)):
        ldi, ldo = m + pad, m + 2 * pad
        x = (rng.standard_normal(ldi * n) * np.exp(rng.uniform(-8, 8, ldi * n))).astype(np.float32)
        plane = ldo * n + 5
        for name, nplanes in (("UNZIP", 2), ("DECOMP_FP32_TO_BF16X2", 2), ("DECOMP_FP32_TO_BF16X3", 3)):
            op = getattr(X, "MELTW_TYPE_UNARY_" + name)
            o0 = rng.integers(0, 60000, size=plane * nplanes, dtype=np.uint16)
            offs = np.array([plane * 2, plane * 4], dtype=np.uint64)
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, gen.F32, gen.BF16, gen.F32), 0)
            assert k, name
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary, p.out.secondary = d_x.data_ptr(), d_o.data_ptr(), offs.ctypes.data
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary, q.out.secondary = x.ctypes.data, want.ctypes.data, offs.ctypes.data
            _ref_call(_desc(1, op, 0, m, n, ldi, 0, 0, ldo, gen.F32, UNS, UNS, gen.BF16, gen.F32), q)
            assert np.array_equal(host(d_o, np.uint16), want), (name, m, n)


def test_large_shapes_take_the_bandwidth_kernels_and_stay_exact():
    """sizes at which the tiled transpose, the vectorised VNNI packers and the two-phase column reduction are selected; data movement
    bit-exact against the oracle restatement, the reduction within the summation-order tolerance"""
    rng = np.random.default_rng(61)
    # transposes with ragged edges and padded leading dimensions
    for t in (gen.F64, gen.F32, gen.BF16, gen.I8):
        for (m, n, pi, po) in ((1000, 777, 0, 0), (257, 129, 3, 5)):
            ldi, ldo = m + pi, n + po
            x = rng.integers(0, 256, size=ldi * n * gen.TS[t], dtype=np.uint8); o0 = rng.integers(0, 256, size=ldo * m * gen.TS[t], dtype=np.uint8)
            op = X.MELTW_TYPE_UNARY_TRANSFORM_NORM_TO_NORMT
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, t), 0)
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
            assert oracle["meltw"](iarr(*_desc(1, op, 0, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, t)), C.addressof(q), 0) == 0
            assert np.array_equal(host(d_o, np.uint8), want), ("normt", t, m, n)
    # VNNI packers
    for name, t in (("NORM_TO_VNNI2", gen.BF16), ("NORM_TO_VNNI2_PAD", gen.BF16), ("NORM_TO_VNNI4", gen.I8), ("NORM_TO_VNNI4_PAD", gen.I8)):
        op = getattr(X, "MELTW_TYPE_UNARY_TRANSFORM_" + name)
        for (m, n, pad) in ((512, 300, 0), (130, 64, 6)):
            ldi, ldo = m + pad, m + pad
            x = rng.integers(0, 256, size=(ldi + 8) * (n + 8) * gen.TS[t], dtype=np.uint8)
            o0 = rng.integers(0, 256, size=(ldo + 8) * (n + 8) * gen.TS[t], dtype=np.uint8)
            k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, t, t, t), 0)
            d_x, d_o = dev(x), dev(o0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            want = o0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
            assert oracle["meltw"](iarr(*_desc(1, op, 0, m, n, ldi, 0, 0, ldo, t, UNS, UNS, t, t)), C.addressof(q), 0) == 0
            assert np.array_equal(host(d_o, np.uint8), want), (name, m, n, pad)
    # column reductions (one result per row): sums and sums of squares
    for t in (gen.F32, gen.BF16):
        for opname in ("REDUCE_X_OP_ADD", "REDUCE_X2_OP_ADD", "REDUCE_X_X2_OP_ADD"):
            for init_acc in (0, X.MELTW_FLAG_UNARY_REDUCE_INIT_ACC):
                m, n, ld = 2048, 600, 2050
                op = getattr(X, "MELTW_TYPE_UNARY_" + opname)
                flags = X.MELTW_FLAG_UNARY_REDUCE_COLS | init_acc
                x = _rand(rng, ld * n, t); y0 = _rand(rng, 2 * ld, t)
                k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, t, t, gen.F32), flags)
                assert k
                d_x, d_y = dev(x), dev(y0)
                p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_y.data_ptr()
                X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
                want = y0.copy(); q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
                assert oracle["meltw"](iarr(*_desc(1, op, flags, m, n, ld, 0, 0, ld, t, UNS, UNS, t, gen.F32)), C.addressof(q), 0) == 0
                _cmp(host(d_y, want.dtype), want, t, False, 2e-2 if t == gen.BF16 else 2e-4)


def test_block_scaled_quantisers_bit_exact():
    """bf16 -> MXFP4 / NVFP4 / MXBF8 (reference :1896-2073, :2247-2326): data bytes and scale bytes, device buffers and host buffers"""
    from test_oracle_meltw import mx_inputs
    rng = np.random.default_rng(71)
    for tout, blk in ((gen.MXFP4X2, 32), (gen.NVFP4X2, 16), (gen.MXBF8, 32)):
        for (m, n, ldi, ldo) in ((64, 9, 64, 64), (96, 8, 100, 128), (32, 7, 32, 32), (512, 300, 512, 512)):
            x = mx_inputs(rng, m, n, ldi)
            y0 = rng.integers(0, 255, size=ldo * n, dtype=np.uint8); s0 = rng.integers(0, 255, size=(ldo // blk) * n + 8, dtype=np.uint8)
            k = X.libxsmm_dispatch_meltw_unary(X.MELTW_TYPE_UNARY_QUANT, X.libxsmm_create_meltw_unary_shape(m, n, ldi, ldo, gen.BF16, tout, gen.F32), 0)
            assert k, (tout, m, n)
            want, wscl = y0.copy(), s0.copy()
            q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary, q.out.secondary = x.ctypes.data, want.ctypes.data, wscl.ctypes.data
            _ref_call(_desc(1, X.MELTW_TYPE_UNARY_QUANT, 0, m, n, ldi, 0, 0, ldo, gen.BF16, UNS, UNS, tout, gen.F32), q)
            d_x, d_y, d_s = dev(x), dev(y0), dev(s0)
            p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary, p.out.secondary = d_x.data_ptr(), d_y.data_ptr(), d_s.data_ptr()
            X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
            assert np.array_equal(host(d_y, np.uint8), want), (tout, m, n, "data")
            assert np.array_equal(host(d_s, np.uint8), wscl), (tout, m, n, "scales")
            if m == 96:      # plain host buffers: staged through the device arena with the exact byte extents
                hy, hs = y0.copy(), s0.copy()
                p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary, p.out.secondary = x.ctypes.data, hy.ctypes.data, hs.ctypes.data
                X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
                assert np.array_equal(hy, want) and np.array_equal(hs, wscl), (tout, "host buffers")


def test_eight_bit_floats_stochastic_rounding_and_dump():
    """BF8 / HF8 element types, STOCHASTIC_ROUND to BF8 (bytes and the advanced generator state) and DUMP against the reference"""
    rng = np.random.default_rng(73)
    m, n, ld = 70, 19, 72
    wide = (rng.standard_normal(ld * n) * np.exp2(rng.integers(-20, 16, size=ld * n))).astype(np.float32)
    wide[3] = np.inf; wide[7] = np.nan; wide[9] = 448.0; wide[10] = 464.0; wide[12] = 57344.0; wide[13] = 61440.0
    allbytes = np.resize(np.arange(256, dtype=np.uint8), ld * n)
    nbytes = {gen.F32: 4, gen.BF16: 2, gen.F16: 2, gen.BF8: 1, gen.HF8: 1}
    for t8 in (gen.BF8, gen.HF8):
        for name in ("IDENTITY", "X2", "RELU"):
            op = getattr(X, "MELTW_TYPE_UNARY_" + name)
            for tin, tout, x in ((gen.F32, t8, wide), (t8, gen.F32, allbytes), (t8, t8, allbytes), (t8, gen.BF16, allbytes)):
                if name == "X2":      # arithmetic on a NaN: x86 keeps the payload, the GPU returns the canonical NaN -- not a rounding question
                    x = np.where(np.isnan(wide), np.float32(1.5), wide) if tin == gen.F32 else np.where((allbytes & 0x7f) > (0x7c if t8 == gen.BF8 else 0x7e), np.uint8(0x3c), allbytes)
                k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, tin, tout, gen.F32), 0)
                assert k, (name, tin, tout)
                want = np.zeros(ld * n * nbytes[tout], dtype=np.uint8)
                q = X.MeltwUnaryParam(); q.inp.primary, q.out.primary = x.ctypes.data, want.ctypes.data
                _ref_call(_desc(1, op, 0, m, n, ld, 0, 0, ld, tin, UNS, UNS, tout, gen.F32), q)
                d_x, d_o = dev(x), dev(np.zeros_like(want))
                p = X.MeltwUnaryParam(); p.inp.primary, p.out.primary = d_x.data_ptr(), d_o.data_ptr()
                X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
                got = host(d_o, np.uint8)
                bad = np.nonzero(got != want)[0]
                assert bad.size == 0, (name, tin, tout, bad[:8].tolist(), got[bad[:8]].tolist(), want[bad[:8]].tolist(), (x.view(np.uint32) if tin == gen.F32 else x)[bad[:8] // nbytes[tout]].tolist())
    # stochastic rounding: unary, binary, ternary (no NaN among the inputs: arithmetic on a NaN keeps the payload on x86 only)
    wide = np.where(np.isnan(wide), np.float32(2.5), wide).astype(np.float32)
    y = rng.standard_normal(ld * n).astype(np.float32); z = rng.standard_normal(ld * n).astype(np.float32)
    state0 = rng.integers(0, 2 ** 32, size=64, dtype=np.uint32)
    for name in ("IDENTITY", "X2", "DUMP"):
        op = getattr(X, "MELTW_TYPE_UNARY_" + name)
        k = X.libxsmm_dispatch_meltw_unary(op, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, gen.F32, gen.BF8, gen.F32), X.MELTW_FLAG_UNARY_STOCHASTIC_ROUND)
        assert k, name
        want, wst, wdump = np.zeros(ld * n, dtype=np.uint8), state0.copy(), np.zeros(ld * n, dtype=np.uint8)
        q = X.MeltwUnaryParam(); q.op.secondary = wst.ctypes.data; q.inp.primary, q.out.primary, q.out.secondary = wide.ctypes.data, want.ctypes.data, wdump.ctypes.data
        _ref_call(_desc(1, op, X.MELTW_FLAG_UNARY_STOCHASTIC_ROUND, m, n, ld, 0, 0, ld, gen.F32, UNS, UNS, gen.BF8, gen.F32), q)
        d_x, d_o, d_s, d_d = dev(wide), dev(np.zeros(ld * n, dtype=np.uint8)), dev(state0), dev(np.zeros(ld * n, dtype=np.uint8))
        p = X.MeltwUnaryParam(); p.op.secondary = d_s.data_ptr(); p.inp.primary, p.out.primary, p.out.secondary = d_x.data_ptr(), d_o.data_ptr(), d_d.data_ptr()
        X.MELTW_UNARY_FN(k)(C.byref(p)); X.check()
        assert np.array_equal(host(d_o, np.uint8), want), name
        assert np.array_equal(host(d_s, np.uint32), wst), (name, "generator state")
        if name == "DUMP":
            assert np.array_equal(host(d_d, np.uint8), wdump)
    kb = X.libxsmm_dispatch_meltw_binary(X.MELTW_TYPE_BINARY_MUL, X.libxsmm_create_meltw_binary_shape(m, n, ld, ld, ld, gen.F32, gen.F32, gen.BF8, gen.F32), X.MELTW_FLAG_BINARY_STOCHASTIC_ROUND)
    assert kb
    want, wst = np.zeros(ld * n, dtype=np.uint8), state0.copy()
    q = X.MeltwBinaryParam(); q.op.secondary = wst.ctypes.data; q.in0.primary, q.in1.primary, q.out.primary = wide.ctypes.data, y.ctypes.data, want.ctypes.data
    _ref_call(_desc(2, X.MELTW_TYPE_BINARY_MUL, X.MELTW_FLAG_BINARY_STOCHASTIC_ROUND, m, n, ld, ld, 0, ld, gen.F32, gen.F32, UNS, gen.BF8, gen.F32), q)
    hs, ho = state0.copy(), np.zeros(ld * n, dtype=np.uint8)       # host buffers all the way: state staged in and out
    p = X.MeltwBinaryParam(); p.op.secondary = hs.ctypes.data; p.in0.primary, p.in1.primary, p.out.primary = wide.ctypes.data, y.ctypes.data, ho.ctypes.data
    X.MELTW_BINARY_FN(kb)(C.byref(p)); X.check()
    assert np.array_equal(ho, want) and np.array_equal(hs, wst)
    kt = X.libxsmm_dispatch_meltw_ternary(X.MELTW_TYPE_TERNARY_MULADD, X.libxsmm_create_meltw_ternary_shape(m, n, ld, ld, ld, ld, gen.F32, gen.F32, gen.F32, gen.BF8, gen.F32), X.MELTW_FLAG_TERNARY_STOCHASTIC_ROUND)
    assert kt
    want, wst = np.zeros(ld * n, dtype=np.uint8), state0.copy()
    q = X.MeltwTernaryParam(); q.op.secondary = wst.ctypes.data; q.in0.primary, q.in1.primary, q.in2.primary, q.out.primary = wide.ctypes.data, y.ctypes.data, z.ctypes.data, want.ctypes.data
    _ref_call(_desc(3, X.MELTW_TYPE_TERNARY_MULADD, X.MELTW_FLAG_TERNARY_STOCHASTIC_ROUND, m, n, ld, ld, ld, ld, gen.F32, gen.F32, gen.F32, gen.BF8, gen.F32), q)
    d_o, d_s = dev(np.zeros(ld * n, dtype=np.uint8)), dev(state0)
    d_x, d_y, d_z = dev(wide), dev(y), dev(z)
    p = X.MeltwTernaryParam(); p.op.secondary = d_s.data_ptr(); p.in0.primary, p.in1.primary, p.in2.primary, p.out.primary = d_x.data_ptr(), d_y.data_ptr(), d_z.data_ptr(), d_o.data_ptr()
    X.MELTW_TERNARY_FN(kt)(C.byref(p)); X.check()
    assert np.array_equal(host(d_o, np.uint8), want) and np.array_equal(host(d_s, np.uint32), wst)
    assert not X.libxsmm_dispatch_meltw_unary(X.MELTW_TYPE_UNARY_DUMP, X.libxsmm_create_meltw_unary_shape(m, n, ld, ld, gen.F64, gen.F64, gen.F64), 0)
