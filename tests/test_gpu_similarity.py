"""-m gpu: CUDA pairwise_distance_matrix (through the C ABI) vs the reference's outputs
(tests/golden/sim_modes.npz) and the oracle."""
import numpy as np
import pytest
import torch

from oracle import similarity as osim

pytestmark = pytest.mark.gpu

MODES = ("fro", "nfro", "euc", "neuc", "sqeuc", "nsqeuc", "cos", "cossim", "dot", "dotsim")


def _wt():
    from wealy_b200 import tensor_ops as wt
    return wt


def _check_mode(got, ref, x, y, mode, single_pass=False):
    """Tolerances: fp16x3 is fp32-grade (|err| ~ 1e-6 on unit vectors); everything is scaled by
    the row norms for the un-normalised modes.  Euclidean modes are compared on squared values
    (sqrt amplifies the cancellation error of |x|^2 - 2xy + |y|^2 near zero, as in the reference)."""
    got, ref = got.double().cpu(), ref.double()
    nx = x.double().norm(dim=-1)[:, None] if x.ndim == 2 else x.double().abs()[:, None]
    ny = y.double().norm(dim=-1)[None, :] if y.ndim == 2 else y.double().abs()[None, :]
    # single pass: every element is rounded to fp16 (2^-11 relative), so short rows see ~1e-3 of |x||y|
    unit = 1.2e-3 if single_pass else 4e-6
    d = x.shape[-1] if x.ndim == 2 else 1
    if mode in ("cos", "cossim"):
        assert (got - ref).abs().max() <= unit
    elif mode in ("dot", "dotsim"):
        assert ((got - ref).abs() <= unit * (nx * ny) + 1e-6).all()
    else:
        scale = {"sqeuc": 1.0, "nsqeuc": float(d), "fro": 1.0, "euc": 1.0, "nfro": d ** 0.5, "neuc": d ** 0.5}[mode]
        g2, r2 = got * scale, ref * scale
        if mode not in ("sqeuc", "nsqeuc"):
            g2, r2 = g2 ** 2, r2 ** 2
        assert ((g2 - r2).abs() <= 4 * unit * (nx ** 2 + ny ** 2) + 1e-6).all()


@pytest.mark.parametrize("case", ["a", "b", "c", "vec"])
@pytest.mark.parametrize("mode", MODES)
def test_modes_against_reference_outputs(golden, case, mode):
    G = golden("sim_modes.npz")
    x, y = torch.from_numpy(G[f"{case}_x"]), torch.from_numpy(G[f"{case}_y"])
    got = _wt().pairwise_distance_matrix(x.cuda(), y.cuda(), mode=mode)
    ref = torch.from_numpy(G[f"{case}_{mode}"])
    assert got.shape == ref.shape and got.dtype == torch.float32 and got.is_cuda
    assert not torch.isnan(got).any()
    _check_mode(got, ref, x, y, mode)


def test_zero_vector_and_duplicate(golden):
    G = golden("sim_modes.npz")
    x, y = torch.from_numpy(G["a_x"]).cuda(), torch.from_numpy(G["a_y"]).cuda()
    cs = _wt().pairwise_distance_matrix(x, y, mode="cossim").cpu()
    assert torch.all(cs[3] == 0)                        # zero vector -> exactly 0, never NaN
    assert abs(float(cs[7, 5]) - 1.0) < 4e-6           # exact duplicate


@pytest.mark.parametrize("tag,dt,tol", [("bf16", torch.bfloat16, 1.6e-2), ("f16", torch.float16, 2e-3)])
def test_output_dtype_follows_input(golden, tag, dt, tol):
    G = golden("sim_modes.npz")
    x = torch.from_numpy(G["dt_x"]).to(dt).cuda()
    got = _wt().pairwise_distance_matrix(x, x, mode="cossim")
    assert got.dtype == dt
    # the reference rounds every intermediate to the half type; we accumulate in fp32 -> compare
    # against both the reference's output and the exact value with the half type's resolution
    exact = osim.distance_matrix(x.cpu().double(), x.cpu().double(), mode="cossim")
    assert (got.cpu().double() - exact).abs().max() <= tol / 2
    assert (got.cpu().double() - torch.from_numpy(G[f"dt_{tag}"])).abs().max() <= tol


@pytest.mark.parametrize("mode", MODES)
def test_float64_inputs_return_float64(golden, mode):
    """The reference returns its input dtype: float64 operands go through the double-precision
    kernel and agree with the reference's own float32 outputs to float32 resolution and with the float64 oracle to 1e-12."""
    G = golden("sim_modes.npz")
    x, y = torch.from_numpy(G["b_x"]).double(), torch.from_numpy(G["b_y"]).double()
    got = _wt().pairwise_distance_matrix(x.cuda(), y.cuda(), mode=mode)
    assert got.dtype == torch.float64 and got.is_cuda
    ref64 = osim.distance_matrix(x, y, mode=mode)
    scale = 1.0 + float(ref64.abs().max())
    assert float((got.cpu() - ref64).abs().max()) <= 1e-12 * scale
    _check_mode(got, torch.from_numpy(G[f"b_{mode}"]), x, y, mode)
    g = torch.Generator().manual_seed(5)
    big_x, big_y = torch.randn(130, 77, generator=g).double(), torch.randn(67, 77, generator=g).double()
    got = _wt().pairwise_distance_matrix(big_x.cuda(), big_y.cuda(), mode=mode)           # ragged tiles of the DGEMM
    ref64 = osim.distance_matrix(big_x, big_y, mode=mode)
    assert float((got.cpu() - ref64).abs().max()) <= 1e-12 * (1.0 + float(ref64.abs().max()))


@pytest.mark.parametrize("mode", MODES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.float64, torch.bfloat16])
def test_every_mode_is_differentiable(mode, dtype):
    """lib/tensor_ops.py:131-176 is plain torch, hence differentiable in every mode: gradients against autograd of the
    oracle restatement in float64 -- including pairs at zero distance (duplicate rows), where the reference's masked
    safe sqrt and cdist's backward both give a zero gradient."""
    wt = _wt()
    g = torch.Generator().manual_seed(17)
    x = (torch.randn(90, 48, generator=g) * 1.5).to(dtype)
    y = torch.randn(70, 48, generator=g).to(dtype)
    y[3] = x[5]                                            # a zero-distance pair
    w = torch.randn(90, 70, generator=g)
    xr, yr = x.double().clone().requires_grad_(True), y.double().clone().requires_grad_(True)
    (osim.distance_matrix(xr, yr, mode=mode) * w.double()).sum().backward()
    xc, yc = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    out = wt.pairwise_distance_matrix(xc, yc, mode=mode)
    assert out.dtype == dtype
    (out * w.cuda().to(dtype)).sum().backward()
    tol = {torch.float32: 5e-5, torch.float64: 1e-10, torch.bfloat16: 3e-2}[dtype]
    assert torch.isfinite(xc.grad).all() and torch.isfinite(yc.grad).all()
    # the rows of the zero-distance pair are compared apart: torch.cdist's matmul path returns 2.4e-7 for that pair in
    # float64 and its backward then adds a spurious O(1e-8) term of arbitrary direction (a direct evaluation of
    # sum_j w_ij (x_i - y_j) / |x_i - y_j| agrees with the CUDA path to 2e-14)
    kx, ky = torch.ones(90, dtype=torch.bool), torch.ones(70, dtype=torch.bool)
    kx[5], ky[3] = False, False
    gx, gy = xc.grad.double().cpu(), yc.grad.double().cpu()
    assert float((gx[kx] - xr.grad[kx]).norm()) <= tol * float(xr.grad.norm())
    assert float((gy[ky] - yr.grad[ky]).norm()) <= tol * float(yr.grad.norm())
    assert float((gx[5] - xr.grad[5]).norm()) <= max(tol, 1e-6) * float(xr.grad[5].norm())
    assert float((gy[3] - yr.grad[3]).norm()) <= max(tol, 1e-6) * float(yr.grad[3].norm())


def test_euclidean_function_is_differentiable():
    wt = _wt()
    g = torch.Generator().manual_seed(19)
    x, y = torch.randn(40, 32, generator=g), torch.randn(55, 32, generator=g)
    for squared in (True, False):
        xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
        osim.euclidean(xr, yr, squared=squared).sum().backward()
        xc, yc = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
        wt.pairwise_euclidean_distance_matrix(xc, yc, squared=squared).sum().backward()
        assert float((xc.grad.double().cpu() - xr.grad).norm()) <= 5e-5 * float(xr.grad.norm())
        assert float((yc.grad.double().cpu() - yr.grad).norm()) <= 5e-5 * float(yr.grad.norm())


@pytest.mark.parametrize("n,m,d", [(1, 1, 1), (1, 300, 7), (129, 255, 65), (257, 513, 1024), (128, 256, 64), (5, 3, 2000)])
@pytest.mark.parametrize("precision", ["fp16x3", "fp16"])
def test_ragged_shapes_against_oracle(n, m, d, precision):
    g = torch.Generator().manual_seed(n * 1000 + m + d)
    x = torch.randn(n, d, generator=g) * 2
    y = torch.randn(m, d, generator=g) * 0.5
    for mode in ("cossim", "dot", "sqeuc"):
        ref = osim.distance_matrix(x.double(), y.double(), mode=mode)
        got = _wt().pairwise_distance_matrix(x.cuda(), y.cuda(), mode=mode, precision=precision)
        _check_mode(got, ref, x, y, mode, single_pass=(precision == "fp16"))


def test_empty_inputs():
    x = torch.randn(0, 16, device="cuda")
    y = torch.randn(5, 16, device="cuda")
    assert _wt().pairwise_distance_matrix(x, y, mode="cos").shape == (0, 5)
    assert _wt().pairwise_distance_matrix(y, x, mode="cos").shape == (5, 0)


def test_row_strided_input_and_same_tensor_fast_path():
    g = torch.Generator().manual_seed(3)
    big = torch.randn(200, 96, generator=g).cuda()
    x = big[:, :40]                                    # row stride 96, inner stride 1
    ref = osim.distance_matrix(x.cpu().double(), x.cpu().double(), mode="cos")
    got = _wt().pairwise_distance_matrix(x, x, mode="cos")
    assert (got.cpu().double() - ref).abs().max() <= 4e-6
    xt = big.t()[:50]                                  # inner stride != 1 -> made contiguous
    ref = osim.distance_matrix(xt.cpu().double(), xt.cpu().double(), mode="cossim")
    got = _wt().pairwise_distance_matrix(xt, xt, mode="cossim")
    assert (got.cpu().double() - ref).abs().max() <= 4e-6


def test_metamorphic_identities():
    g = torch.Generator().manual_seed(11)
    x = torch.randn(300, 128, generator=g).cuda()
    y = torch.randn(200, 128, generator=g).cuda()
    wt = _wt()
    cs = wt.pairwise_distance_matrix(x, y, mode="cossim")
    assert torch.equal(wt.pairwise_distance_matrix(x, y, mode="cos"), 1 - cs)          # cos == 1 - cossim
    scaled = wt.pairwise_distance_matrix(x * 8.0, y * 0.25, mode="cossim")             # power-of-two scale invariance
    assert (scaled - cs).abs().max() <= 2e-6
    assert (wt.pairwise_distance_matrix(y, x, mode="cossim").t() - cs).abs().max() <= 2e-6   # symmetry
    eu = wt.pairwise_euclidean_distance_matrix(x, y)
    sq = wt.pairwise_euclidean_distance_matrix(x, y, squared=True)
    assert (eu ** 2 - sq).abs().max() <= 1e-3


def test_error_behaviour_matches_reference():
    x = torch.randn(4, 8, device="cuda")
    with pytest.raises(NotImplementedError):
        _wt().pairwise_distance_matrix(x, x, mode="nope")          # lib/tensor_ops.py:175
    with pytest.raises(NotImplementedError):
        _wt().pairwise_distance_matrix(x, x, mode="fro", p=3)      # documented gap: cdist p != 2
    with pytest.raises(AssertionError):
        _wt().pairwise_distance_matrix(x[None], x[None])            # lib/tensor_ops.py:153


@pytest.mark.parametrize("mode", ["cossim", "cos", "dotsim", "dot"])
@pytest.mark.parametrize("n,m,d,dtype", [(70, 130, 96, torch.float32), (300, 300, 1024, torch.float32),
                                         (129, 65, 200, torch.bfloat16)])
def test_cosine_modes_are_differentiable(mode, n, m, d, dtype):
    """The reference differentiates through pairwise_distance_matrix (lib/losses.py:45): gradients of the cosine
    modes against autograd of the oracle restatement (fp64)."""
    from wealy_b200 import tensor_ops as wt
    from oracle import similarity as osim
    g = torch.Generator().manual_seed(n + m)
    x = (torch.randn(n, d, generator=g) * 2).to(dtype)
    y = torch.randn(m, d, generator=g).to(dtype)
    w = torch.randn(n, m, generator=g)
    xr, yr = x.double().requires_grad_(True), y.double().requires_grad_(True)
    (osim.distance_matrix(xr, yr, mode=mode) * w.double()).sum().backward()
    xc, yc = x.cuda().requires_grad_(True), y.cuda().requires_grad_(True)
    out = wt.pairwise_distance_matrix(xc, yc, mode=mode)
    (out * w.cuda().to(dtype)).sum().backward()
    tol = 2e-5 if dtype == torch.float32 else 2e-2
    assert float((xc.grad.double().cpu() - xr.grad).norm()) <= tol * float(xr.grad.norm())
    assert float((yc.grad.double().cpu() - yr.grad).norm()) <= tol * float(yr.grad.norm())
    # the same tensor on both sides: the two contributions add up
    xs = x.cuda().requires_grad_(True)
    (wt.pairwise_distance_matrix(xs, xs, mode=mode)[:, :n] * w[:, :n].cuda().to(dtype) if m >= n else
     wt.pairwise_distance_matrix(xs, xs, mode=mode)).sum().backward()
    assert torch.isfinite(xs.grad).all()
