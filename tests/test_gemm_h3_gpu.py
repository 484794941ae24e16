"""The scaled 3xFP16 tensor-core contraction behind every GSM GEMM (gsmvi/gsm.py:11-27, 53-54, 119) against an fp64
product of the same (dequantised) operands: both kernels behind gsmvi_gemm_h3 - the persistent 2-CTA kernel
(h3x2_gemm.cuh) and the one-CTA kernel (h3_gemm.cuh) - over ragged shapes, both operand layouts, the triangular K range
of the sampler, lower-triangle + mirror output, beta / bias, and the |C| max the next split consumes."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from gsmvi_b200 import _lib
    _lib.lib()
    prev = _lib.h3_pair_kernel(-1)
    yield _lib
    _lib.h3_pair_kernel(prev)


def _record(row):
    import json, os
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_gemm_h3.jsonl"), "a") as f:
        f.write(json.dumps(row) + "\n")


def _operand(L, A):
    t = torch.from_numpy(np.ascontiguousarray(A, dtype=np.float32)).cuda()
    return L.HOperand(t.shape[0], t.shape[1], "cuda").split_from(t)


def _ref(h):
    return h.dequant().double().cpu().numpy()


CASES = [
    # M, N, K, a_mn, b_mn, opts
    (256, 128, 64, False, False, {}),
    (512, 384, 1024, False, False, {}),
    (384, 256, 1100, False, False, {"alpha": -0.5, "bias": True}),          # odd tile row count, ragged K, two chunks
    (1000, 520, 333, False, False, {"beta": 0.75}),                          # ragged everything
    (640, 640, 640, False, False, {"krange_b_lower": True, "bias": True}),   # sampler: B = L lower triangular
    (2048, 1024, 2304, False, False, {}),                                    # three chunks, several tiles per pair
    (512, 512, 768, True, True, {}),                                         # MN-major views (the statistics GEMM)
    (640, 640, 1536, True, True, {"tri": True, "mirror": True, "beta": 1.0, "alpha": -1.0 / 768}),   # covariance update
    (1000, 1000, 300, True, True, {"tri": True, "mirror": True}),            # ragged tri, odd tile count
    (768, 512, 512, True, False, {}),
    (512, 768, 512, False, True, {"bias": True}),
    (4096, 256, 2048, False, False, {}),                                     # more supertiles than pairs: snake rounds
]


@pytest.mark.parametrize("pair", [1, 0])
@pytest.mark.parametrize("M,N,K,a_mn,b_mn,opts", CASES)
def test_gemm_h3_matches_fp64(L, pair, M, N, K, a_mn, b_mn, opts):
    rng = np.random.RandomState(M * 7 + N * 3 + K + 2 * a_mn + b_mn)
    A = rng.normal(size=(K, M) if a_mn else (M, K))
    B = rng.normal(size=(K, N) if b_mn else (N, K))
    if opts.get("krange_b_lower"):
        B = np.tril(B)
    Ah, Bh = _operand(L, A), _operand(L, B)
    Ar, Br = _ref(Ah), _ref(Bh)
    Ar = Ar.T if a_mn else Ar
    Br = Br.T if b_mn else Br
    alpha, beta = opts.get("alpha", 1.0), opts.get("beta", 0.0)
    bias = torch.from_numpy(rng.normal(size=N).astype(np.float32)).cuda() if opts.get("bias") else None
    ldc = (N + 31) // 32 * 32
    Cin = None
    if beta != 0.0:
        c0 = rng.normal(size=(M, N))
        if opts.get("tri"):
            c0 = (c0 + c0.T) / 2
        Cin = torch.zeros(M, ldc, device="cuda")
        Cin[:, :N] = torch.from_numpy(c0.astype(np.float32)).cuda()
    want = alpha * (Ar @ Br.T)
    if Cin is not None:
        want = want + beta * Cin[:, :N].double().cpu().numpy()
    if bias is not None:
        want = want + bias.double().cpu().numpy()[None, :]
    C = torch.full((M, ldc), float("nan"), device="cuda")
    amax = torch.zeros(1, dtype=torch.int32, device="cuda")
    L.h3_pair_kernel(pair)
    L.gemm_h3(Ah, Bh, C, M, N, K, a_mn=a_mn, b_mn=b_mn, alpha=alpha, beta=beta, Cin=Cin, bias_n=bias,
              tri=bool(opts.get("tri")), mirror=bool(opts.get("mirror")),
              krange=L.KR_B_LOWER if opts.get("krange_b_lower") else 0, absmax_out=amax)
    torch.cuda.synchronize()
    got = C[:, :N].double().cpu().numpy()
    assert np.isfinite(np.tril(got) if opts.get("tri") else got).all()
    if opts.get("mirror"):
        assert np.array_equal(got, got.T)       # the mirrored stores make the result exactly symmetric
    if opts.get("tri"):                         # lower tiles are computed (and mirrored): the lower triangle is the product
        got, want = np.tril(got), np.tril(want)
    scale = np.abs(alpha) * np.sqrt(K) + 1.0
    err = np.max(np.abs(got - want)) / scale
    _record(dict(M=M, N=N, K=K, a_mn=a_mn, b_mn=b_mn, pair=pair, opts=sorted(opts), max_err_over_sqrtK=float(err),
                 relF=float(np.linalg.norm(got - want) / np.linalg.norm(want))))
    # fp32-grade: 22-bit products; TMEM accumulation truncates, so the bar leaves room for the one-CTA kernel's K/3-long
    # accumulators; the pair kernel's 32-addition chunks stay at ~4e-6 whatever K is
    assert err < (6e-6 if pair else 8e-6), err
    got_max = amax.view(torch.float32).item()
    assert got_max == pytest.approx(np.max(np.abs(got)), rel=1e-6)
    assert torch.isnan(C[:, N:]).all()          # nothing written outside the result


def test_pair_kernel_and_single_kernel_agree_closely(L):
    """Same operands through both kernels: they differ only in the order the partial sums meet (chunked fp32 registers against
    round-robin TMEM accumulators), i.e. at the 1e-7 level."""
    rng = np.random.RandomState(5)
    M = N = K = 1024
    Ah, Bh = _operand(L, rng.normal(size=(M, K))), _operand(L, rng.normal(size=(N, K)))
    out = []
    for pair in (1, 0):
        L.h3_pair_kernel(pair)
        C = torch.empty(M, N, device="cuda")
        L.gemm_h3(Ah, Bh, C, M, N, K)
        out.append(C.double().cpu().numpy())
    assert np.max(np.abs(out[0] - out[1])) < 8e-6 * np.sqrt(K)
