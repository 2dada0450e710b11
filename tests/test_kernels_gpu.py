"""Kernel-level parity: every C-ABI entry point against the oracle / fp64 torch maths on the same inputs.

All tests call through ``libvqa_sm100.so`` (ctypes) on the GPU; the oracle (``oracle/vqa_oracle.py``) and
plain fp64 torch expressions are only the checkers.  Tolerances are max-norm relative errors
(``conftest.rel_err``); fp32 budget from BASELINE.json north_star: 1e-3.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, golden_params, rel_err
from oracle import vqa_oracle as O

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def kn():
    from vqa_b200 import kernels
    return kernels


def _pack_gauss(p, prefix, device=DEV, dtype=torch.float32):
    return torch.cat([p[f"{prefix}.{k}"].reshape(-1) for k in
                      ("mean_rho", "precision_rho", "mean_theta", "precision_theta")]).to(device=device, dtype=dtype).contiguous()


# --------------------------------------------------------------------------------------------- GEMM
GEMM_CASES = [
    # M, N, K, a_mn, b_mn, tile_n
    (128, 256, 64, False, False, 0),
    (300, 520, 132, False, False, 0),      # ragged M/N/K tails (TMA zero fill + predicated epilogue)
    (256, 512, 2052, False, False, 256),   # F = 2052: K not a multiple of the 32-wide k-block
    (256, 128, 96, False, False, 128),
    (64, 64, 32, False, False, 64),
    (200, 264, 160, False, True, 0),       # dX = dY . W : B stored (K, N)
    (264, 200, 300, True, True, 0),        # dW = dY^T . X : both operands MN-major
    (128, 3000, 512, True, True, 256),
    (36, 512, 3076, False, False, 0),      # GraphLearner layer-1 shape on one image
]


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,tile_n", GEMM_CASES)
@pytest.mark.parametrize("prec,tol", [(0, 5e-5), (1, 3e-3), (2, 3e-6)])  # x3: tensor-core accumulators truncate (~K*2^-24)
def test_gemm_matches_fp64(kn, M, N, K, a_mn, b_mn, tile_n, prec, tol):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g)
    ref = a.double() @ b.double().t()
    a_dev = (a.t().contiguous() if a_mn else a).to(DEV)
    b_dev = (b.t().contiguous() if b_mn else b).to(DEV)
    out = kn.gemm(a_dev, b_dev, a_mn=a_mn, b_mn=b_mn, precision=prec, tile_n=tile_n)
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    assert rel_err(out.cpu(), ref) < tol


def test_gemm_epilogues_and_views(kn):
    g = torch.Generator().manual_seed(5)
    B, Kn, F, N = 5, 12, 40, 72
    M = B * Kn
    x = torch.randn(M, F, generator=g)
    wfull = torch.randn(N, F + 16, generator=g)          # weight with extra columns: use a column sub-view (ld != K)
    w = wfull[:, :F]
    bias = torch.randn(N, generator=g)
    rb = torch.randn(B, N, generator=g)
    aux = torch.randn(M, N, generator=g)
    ref = x.double() @ w.double().t() + rb.double().repeat_interleave(Kn, 0) + bias.double()
    ref_relu = ref.clamp(min=0)
    out = kn.gemm(x.to(DEV), wfull.to(DEV)[:, :F], bias=bias.to(DEV), rowbcast=rb.to(DEV), group=Kn, relu=True)
    assert rel_err(out.cpu(), ref_relu) < 5e-5
    # aux mask + scale (ReLU / dropout backward fused into the dX GEMM)
    out2 = kn.gemm(x.to(DEV), w.contiguous().to(DEV), aux=aux.to(DEV), aux_scale=2.0)
    ref2 = torch.where(aux > 0, 2.0 * (x.double() @ w.double().t()), torch.zeros((), dtype=torch.float64))
    assert rel_err(out2.cpu(), ref2) < 5e-5
    # write into a column slice of a wider buffer (ldc != N)
    big = torch.zeros(M, N + 24, device=DEV)
    kn.gemm(x.to(DEV), w.contiguous().to(DEV), out=big[:, 8:8 + N])
    assert rel_err(big[:, 8:8 + N].cpu(), x.double() @ w.double().t()) < 5e-5
    assert big[:, :8].abs().max() == 0 and big[:, 8 + N:].abs().max() == 0


@pytest.mark.parametrize("split", [2, 5, 16])
def test_gemm_split_k(kn, split):
    g = torch.Generator().manual_seed(9)
    Kc, M, N = 1000, 96, 200
    a = torch.randn(Kc, M, generator=g)
    b = torch.randn(Kc, N, generator=g)
    out = kn.gemm(a.to(DEV), b.to(DEV), a_mn=True, b_mn=True, split_k=split)
    assert rel_err(out.cpu(), a.double().t() @ b.double()) < 5e-5


def test_gemm_rejects_bad_arguments(kn):
    a = torch.randn(8, 30, device=DEV)     # ld = 30: not a multiple of 4 floats -> TMA illegal
    b = torch.randn(8, 30, device=DEV)
    with pytest.raises(RuntimeError):
        kn.gemm(a, b)
    with pytest.raises(RuntimeError):
        kn.gemm(torch.randn(8, 32), torch.randn(8, 32))   # CPU tensors: no fallback


# --------------------------------------------------------------------------------------------- split-bf16 GEMM
def test_split_planes_reconstruct_fp32(kn):
    x = torch.randn(37, 2052, device=DEV) * 3.0
    s = kn.split(x)
    assert s.ld == 2056 and s.hi.dtype == torch.bfloat16
    assert rel_err(s.float().cpu(), x.cpu()) < 2 ** -16            # hi + lo carries >= 16 mantissa bits
    assert (s.hi[:, :2052].float() == x.bfloat16().float()).all()  # hi is exactly bf16(x), round to nearest
    h = kn.split(x, with_lo=False)
    assert h.lo is None and (h.hi[:, :2052] == s.hi[:, :2052]).all()


def test_dropout_split_matches_dropout_then_split(kn):
    x = torch.randn(37, 2052, device=DEV)
    for p in (0.5, 0.4):
        ref = kn.dropout(x, p, seed=77, offset=5)                    # cols % 4 == 0: identical Philox counters -> identical mask
        s = kn.dropout_split(x, p, seed=77, offset=5)
        assert torch.equal(s.float() != 0, ref != 0)
        assert rel_err(s.float().cpu(), ref.cpu()) < 2 ** -16
    y = torch.randn(11, 30, device=DEV)                              # ragged width: statistics only
    s = kn.dropout_split(y.repeat(200, 1), 0.5, seed=1, offset=1)
    assert abs((s.float() != 0).float().mean().item() - 0.5) < 0.02
    step = torch.tensor(3, dtype=torch.int64, device=DEV)
    assert not torch.equal(kn.dropout_split(x, 0.5, 77, 5, step).hi, kn.dropout_split(x, 0.5, 77, 5).hi)


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,tile_n", GEMM_CASES)
@pytest.mark.parametrize("passes,tol", [(3, 3e-5), (1, 8e-3)])
def test_gemm_split_bf16_matches_fp64(kn, M, N, K, a_mn, b_mn, tile_n, passes, tol):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g)
    ref = a.double() @ b.double().t()
    a_s = kn.split((a.t().contiguous() if a_mn else a).to(DEV))
    b_s = kn.split((b.t().contiguous() if b_mn else b).to(DEV))
    out = kn.gemm_s(a_s, b_s, a_mn=a_mn, b_mn=b_mn, passes=passes, tile_n=tile_n)
    torch.cuda.synchronize()
    assert out.shape == (M, N)
    assert rel_err(out.cpu(), ref) < tol


def test_gemm_split_bf16_epilogues(kn):
    g = torch.Generator().manual_seed(5)
    B, Kn, F, N = 5, 12, 40, 72
    M = B * Kn
    x = torch.randn(M, F, generator=g)
    w = torch.randn(N, F, generator=g)
    bias = torch.randn(N, generator=g)
    rb = torch.randn(B, N, generator=g)
    aux = torch.randn(M, N, generator=g)
    xs, ws = kn.split(x.to(DEV)), kn.split(w.to(DEV))
    ref = x.double() @ w.double().t() + rb.double().repeat_interleave(Kn, 0) + bias.double()
    out = kn.gemm_s(xs, ws, bias=bias.to(DEV), rowbcast=rb.to(DEV), group=Kn, relu=True)
    assert rel_err(out.cpu(), ref.clamp(min=0)) < 3e-5
    ref2 = torch.where(aux > 0, 2.0 * (x.double() @ w.double().t()), torch.zeros((), dtype=torch.float64))
    out2 = kn.gemm_s(xs, ws, aux=aux.to(DEV), aux_scale=2.0)                       # fp32 mask
    assert rel_err(out2.cpu(), ref2) < 3e-5
    out3 = kn.gemm_s(xs, ws, aux=kn.split(aux.to(DEV)), aux_scale=2.0)             # mask = hi plane of a split tensor
    assert rel_err(out3.cpu(), ref2) < 3e-5
    # split-plane output (with and without the fp32 copy) and writing into a column slice
    os_ = kn.empty_split(M, N, DEV)
    out4 = kn.gemm_s(xs, ws, out_split=os_)
    assert rel_err(os_.float().cpu(), x.double() @ w.double().t()) < 3e-5 and rel_err(out4.cpu(), x.double() @ w.double().t()) < 3e-5
    os2 = kn.empty_split(M, N, DEV)
    r = kn.gemm_s(xs, ws, out_split=os2, want_f32=False, relu=True)
    assert r is os2 and rel_err(os2.float().cpu(), (x.double() @ w.double().t()).clamp(min=0)) < 3e-5
    big = torch.zeros(M, N + 24, device=DEV)
    kn.gemm_s(xs, ws, out=big[:, 8:8 + N])
    assert rel_err(big[:, 8:8 + N].cpu(), x.double() @ w.double().t()) < 3e-5
    assert big[:, :8].abs().max() == 0 and big[:, 8 + N:].abs().max() == 0
    # column sub-view of a wider operand (question half of the graph-learner weight)
    wide = torch.randn(N, 64 + F, generator=g)
    wide_s = kn.split(wide.to(DEV))
    out5 = kn.gemm_s(xs, wide_s.cols_slice(64, 64 + F))
    assert rel_err(out5.cpu(), x.double() @ wide[:, 64:].double().t()) < 3e-5


@pytest.mark.parametrize("split", [2, 5, 16])
def test_gemm_split_bf16_split_k(kn, split):
    g = torch.Generator().manual_seed(9)
    Kc, M, N = 1000, 96, 200
    a = torch.randn(Kc, M, generator=g)
    b = torch.randn(Kc, N, generator=g)
    out = kn.gemm_s(kn.split(a.to(DEV)), kn.split(b.to(DEV)), a_mn=True, b_mn=True, split_k=split)
    assert rel_err(out.cpu(), a.double().t() @ b.double()) < 3e-5


@pytest.mark.parametrize("M,N,K,a_mn,b_mn,split", [(1100, 520, 200, False, False, 1), (1088, 512, 320, False, True, 1),
                                                    (1200, 300, 1000, True, True, 3), (2048, 2052, 1100, True, True, 1)])
@pytest.mark.parametrize("passes,tol", [(3, 3e-5), (1, 8e-3)])
def test_gemm_split_bf16_cta_pairs(kn, M, N, K, a_mn, b_mn, split, passes, tol):
    """>= 8 tile rows at BN = 256: CTA pairs with the B tile multicast to both (odd tile-row counts get a padding tile)."""
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g)
    b = torch.randn(N, K, generator=g)
    ref = a.double() @ b.double().t()
    a_s = kn.split((a.t().contiguous() if a_mn else a).to(DEV))
    b_s = kn.split((b.t().contiguous() if b_mn else b).to(DEV))
    out = kn.gemm_s(a_s, b_s, a_mn=a_mn, b_mn=b_mn, passes=passes, tile_n=256, split_k=split)
    assert rel_err(out.cpu(), ref) < tol
    solo = kn.gemm_s(a_s, b_s, a_mn=a_mn, b_mn=b_mn, passes=passes, tile_n=256, split_k=split, cluster=False)
    if split == 1:
        assert torch.equal(out, solo)                    # same arithmetic, only the operand delivery differs
    else:
        assert rel_err(out.cpu(), solo.cpu().double()) < 1e-6


def test_gemm_split_bf16_accumulate_and_auto_tile(kn):
    g = torch.Generator().manual_seed(11)
    M, N, Kc = 512, 1024, 3072            # the GRU backward product: few tiles -> BN = 64 tile, split-K, accumulate into C
    a = torch.randn(M, Kc, generator=g)
    b = torch.randn(Kc, N, generator=g)
    c0 = torch.randn(M, N, generator=g)
    out = c0.to(DEV).clone()
    kn.gemm_s(kn.split(a.to(DEV)), kn.split(b.to(DEV)), b_mn=True, out=out, accumulate=True, split_k=4)
    assert rel_err(out.cpu(), c0.double() + a.double() @ b.double()) < 3e-5
    out1 = c0.to(DEV).clone()
    kn.gemm_s(kn.split(a.to(DEV)), kn.split(b.to(DEV)), b_mn=True, out=out1, accumulate=True)
    assert rel_err(out1.cpu(), c0.double() + a.double() @ b.double()) < 3e-5


# --------------------------------------------------------------------------------------------- question encoder
@pytest.mark.parametrize("B,T,H,V,E,sort", [(9, 7, 64, 50, 300, False), (33, 14, 128, 200, 300, False), (4, 3, 32, 11, 24, False),
                                            (300, 12, 64, 40, 24, True), (260, 9, 32, 40, 24, False)])
@pytest.mark.parametrize("fused", ["seq", "step", "off"])
def test_question_encoder_matches_packed_gru(B, T, H, V, E, sort, fused, monkeypatch):
    """ops.QuestionEncoderFn (padded, masked recurrence on our kernels) vs nn.Embedding + pack_padded_sequence + nn.GRU
    in fp64 on the CPU (the reference's construction, sparse_graph_model.py:117-121), forward and every gradient.
    The B > 128 cases exercise the row-tile gate of the per-step products (sorted as collate_fn does: whole tiles end early;
    unsorted: tiles stay alive as long as any of their rows is)."""
    from torch.nn.utils.rnn import pack_padded_sequence
    from vqa_b200 import ops
    monkeypatch.setattr(ops, "GRU_FUSED_SEQ", "1" if fused == "seq" else "0")     # all steps in one cooperative launch (grid barrier between steps)
    monkeypatch.setattr(ops, "GRU_FUSED", fused == "step")        # one kernel per step (product + cell out of TMEM); "off": two launches per step
    g = torch.Generator().manual_seed(B * 100 + T)
    lens = torch.randint(1, T + 1, (B,), generator=g)
    lens[0] = T
    if sort:
        lens = lens.sort(descending=True).values
    q = torch.zeros(B, T + 5, dtype=torch.int64)
    for b in range(B):
        q[b, :lens[b]] = torch.randint(1, V, (int(lens[b]),), generator=g)
    emb = torch.nn.Embedding(V, E).double()
    gru = torch.nn.GRU(E, H).double()
    R = torch.randn(B, H, generator=g).double()
    packed = pack_padded_sequence(emb(q), lens, batch_first=True, enforce_sorted=False)
    _, hid = gru(packed)
    (hid[0] * R).sum().backward()
    params = [emb.weight, gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0]
    dev_params = [p.detach().float().to(DEV).requires_grad_(True) for p in params]
    out = ops.QuestionEncoderFn.apply(q.to(DEV), lens.to(torch.int32).to(DEV), int(lens.max()), *dev_params)
    (out * R.float().to(DEV)).sum().backward()
    assert rel_err(out.detach().cpu(), hid[0].detach()) < 2e-5
    for name, p, d in zip(["wembed", "w_ih", "w_hh", "b_ih", "b_hh"], params, dev_params):
        assert rel_err(d.grad.cpu(), p.grad) < 1e-4, name


# --------------------------------------------------------------------------------------------- small kernels
def test_dropout_statistics_and_determinism(kn):
    x = torch.ones(1 << 20, device=DEV)
    for p in (0.5, 0.4, 0.1):
        y = kn.dropout(x, p, seed=1234, offset=7)
        keep = (y != 0).float().mean().item()
        assert abs(keep - (1 - p)) < 4e-3
        assert torch.allclose(y[y != 0], torch.full((), 1 / (1 - p), device=DEV))
        assert torch.equal(y, kn.dropout(x, p, seed=1234, offset=7))
        assert not torch.equal(y, kn.dropout(x, p, seed=1234, offset=8))
    y = kn.dropout(torch.randn(1003, device=DEV), 0.0, 1, 1)   # ragged tail, p = 0 is the identity
    assert y.shape == (1003,)


@pytest.mark.parametrize("cols", [133, 3000, 512])          # scalar path and the float4 path
def test_weight_norm_fwd_bwd(kn, cols):
    torch.manual_seed(3)
    v = torch.randn(70, cols, device=DEV, requires_grad=True)
    g = torch.rand(70, 1, device=DEV, requires_grad=True) + 0.5
    g = g.detach().requires_grad_(True)
    w_ref = torch._weight_norm(v, g, 0)
    w = kn.weight_norm_fwd(v.detach(), g.detach())
    assert rel_err(w.cpu(), w_ref.detach().cpu()) < 1e-6
    dw = torch.randn_like(w_ref)
    w_ref.backward(dw)
    dv, dg = kn.weight_norm_bwd(dw, v.detach(), g.detach())
    assert rel_err(dv.cpu(), v.grad.cpu()) < 1e-5
    assert rel_err(dg.cpu(), g.grad.cpu()) < 1e-5


def test_weight_norm_split_matches_weight_norm_then_split(kn):
    g_ = torch.Generator().manual_seed(8)
    for rows, cols, c0, c1 in [(70, 3076, 0, 2052), (70, 3076, 2052, 3076), (33, 512, 0, 512), (9, 24, 4, 17)]:
        v = torch.randn(rows, cols, generator=g_).to(DEV); g = (torch.rand(rows, 1, generator=g_) + 0.5).to(DEV)
        ref = kn.split(kn.weight_norm_fwd(v, g)[:, c0:c1].contiguous())
        got = kn.weight_norm_split(v, g, c0, c1)
        assert got.shape == ref.shape
        n = c1 - c0
        w64 = v.double() * (g.double() / v.double().norm(dim=1, keepdim=True))
        assert rel_err(got.float()[:, :n].cpu(), w64[:, c0:c1].cpu()) < 1e-5 and rel_err(got.float()[:, :n], ref.float()[:, :n]) < 1e-5   # two bf16 planes carry ~17 mantissa bits
        assert torch.all(got.hi[:, n:] == 0) and torch.all(got.lo[:, n:] == 0)      # plane padding stays zero (TMA reads it)
        assert kn.weight_norm_split(v, g, c0, c1, with_lo=False).lo is None


def test_reductions_and_gate(kn):
    torch.manual_seed(4)
    for shape in [(7168, 3072), (1001, 512), (64, 8), (5, 12)]:                    # vector path incl. ragged row counts
        xb = torch.randn(*shape, device=DEV)
        assert rel_err(kn.colsum(xb).cpu(), xb.double().sum(0).cpu()) < 2e-6, shape
    x = torch.randn(36 * 17, 200, device=DEV)
    assert rel_err(kn.colsum(x).cpu(), x.double().sum(0).cpu()) < 1e-6
    assert rel_err(kn.colsum(x[:, 8:72]).cpu(), x[:, 8:72].double().sum(0).cpu()) < 1e-6
    assert rel_err(kn.segment_sum(x, 36).cpu(), x.double().view(17, 36, 200).sum(1).cpu()) < 1e-6
    dhq, q, pooled = torch.randn(3, 50, device=DEV), torch.randn(3, 50, device=DEV), torch.randn(3, 50, device=DEV).clamp(min=0)
    dpooled, dq = kn.gate_bwd(dhq, q, pooled)
    assert torch.allclose(dpooled, torch.where((pooled > 0) & (q > 0), dhq * q, torch.zeros_like(q)))
    assert torch.allclose(dq, torch.where(q > 0, dhq * pooled, torch.zeros_like(q)))


# --------------------------------------------------------------------------------------------- graph learner tail
@pytest.mark.parametrize("name", ["tiny", "small"])
def test_topk_indices_bit_exact_given_reference_adjacency(kn, name):
    """north_star: top-k neighbourhood indices bit-exact when given the reference's adjacency."""
    g = load_golden(name)
    adj = torch.from_numpy(g["out.adjacency"]).to(DEV)
    nb = g["nbr.idx_sorted"].shape[-1]
    idx, alpha = kn.topk_softmax(adj, nb)
    order = idx.long().argsort(-1)
    assert np.array_equal(torch.gather(idx.long(), -1, order).cpu().numpy(), g["nbr.idx_sorted"])
    assert rel_err(torch.gather(alpha, -1, order).cpu(), g["nbr.alpha_sorted"]) < 1e-6
    # emitted in descending value order
    vals = torch.gather(adj, -1, idx.long())
    assert (vals[..., :-1] >= vals[..., 1:]).all()


@pytest.mark.parametrize("B,K,C,nb", [(7, 36, 512, 16), (3, 51, 512, 19), (2, 100, 512, 32), (4, 12, 64, 5), (2, 128, 128, 128), (3, 5, 8, 1)])
def test_adjacency_topk_fwd(kn, B, K, C, nb):
    torch.manual_seed(B * 100 + K)
    h = torch.randn(B, K, C, device=DEV).clamp(min=0)
    adj, idx, alpha = kn.adjacency_topk_fwd(h, nb)
    ref = h.double() @ h.double().transpose(1, 2)
    assert rel_err(adj.cpu(), ref.cpu()) < 5e-5
    assert torch.equal(adj, adj.transpose(1, 2))            # symmetric by construction
    # selection is exact w.r.t. the adjacency the kernel itself produced
    tv, ti = torch.topk(adj, nb, dim=-1)
    assert torch.equal(idx.long().sort(-1).values, ti.sort(-1).values)
    a_ref = torch.softmax(torch.gather(adj, -1, idx.long()).double(), -1)
    assert rel_err(alpha.cpu(), a_ref.cpu()) < 1e-6


def test_topk_tie_rule(kn):
    adj = torch.zeros(1, 6, 6, device=DEV)
    adj[0, :, 2] = 1.0
    adj[0, :, 4] = 1.0
    idx, alpha = kn.topk_softmax(adj, 3)
    # ties: lower index first; any index whose value equals the k-th value is a valid member (SURVEY.md trap 1)
    assert idx[0, 0].tolist() == [2, 4, 0]
    assert torch.allclose(alpha.sum(-1), torch.ones(1, 6, device=DEV))


@pytest.mark.parametrize("B,K,C,nb,with_dadj", [(5, 36, 512, 16, False), (2, 51, 512, 19, True), (2, 100, 256, 32, False), (3, 12, 64, 5, True)])
def test_adjacency_topk_bwd(kn, B, K, C, nb, with_dadj):
    torch.manual_seed(K)
    hpre = torch.randn(B, K, C, device=DEV, dtype=torch.float64) * 0.08   # keep the softmax unsaturated: fp32 cancellation otherwise
    hpre.requires_grad_(True)
    h = torch.relu(hpre)
    adj = h @ h.transpose(1, 2)
    _, idx32, alpha32 = kn.adjacency_topk_fwd(h.detach().float(), nb)
    idx = idx32.long()
    alpha = torch.softmax(torch.gather(adj, -1, idx), -1)
    dalpha = torch.randn(B, K, nb, device=DEV, dtype=torch.float64)
    dadj = torch.randn(B, K, K, device=DEV, dtype=torch.float64) if with_dadj else None
    loss = (alpha * dalpha).sum() + ((adj * dadj).sum() if with_dadj else 0.0)
    (gref,) = torch.autograd.grad(loss, hpre)
    got = kn.adjacency_topk_bwd(h.detach().float(), idx32, alpha32, dalpha.float(), None if dadj is None else dadj.float())
    assert rel_err(got.cpu(), gref.cpu()) < 2e-5


# --------------------------------------------------------------------------------------------- graph convolution
def _gc_reference(Y, idx, alpha, image, gauss_p, prefix, nk):
    """fp64 oracle of the project-first aggregate: gather rows of Y with the oracle's own Gaussian weights."""
    B, K, nb = idx.shape
    cen = O.box_centres(image)
    pseudo = O.gather_pseudo(O.polar_pseudo_coordinates(cen), idx)
    w = O.gaussian_kernel_weights(pseudo, gauss_p, prefix).view(B, K, nb, nk)
    D = Y.shape[-1] // nk
    nbr = O.gather_neighbours(Y.view(B, K, -1), idx).view(B, K, nb, nk, D)
    coef = w if alpha is None else w * alpha.unsqueeze(-1)
    return (coef.unsqueeze(-1) * nbr).sum(2).reshape(B, K, nk * D)


def _gc_inputs(B, K, nb, nk, out_dim, F=12, seed=0, dtype=torch.float64):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(B, K, F, generator=g, dtype=dtype)
    xy1 = torch.rand(B, K, 2, generator=g, dtype=dtype) * 0.7
    image[..., -4:-2] = xy1
    image[..., -2:] = xy1 + torch.rand(B, K, 2, generator=g, dtype=dtype) * 0.25 + 0.05
    Y = torch.randn(B, K, out_dim, generator=g, dtype=dtype)
    idx = torch.stack([torch.stack([torch.randperm(K, generator=g)[:nb] for _ in range(K)]) for _ in range(B)])
    alpha = torch.softmax(torch.randn(B, K, nb, generator=g, dtype=dtype), -1)
    gp = {"gc.mean_rho": torch.rand(nk, 1, generator=g, dtype=dtype), "gc.mean_theta": (torch.rand(nk, 1, generator=g, dtype=dtype) * 2 - 1) * math.pi,
          "gc.precision_rho": torch.rand(nk, 1, generator=g, dtype=dtype) * 0.9 + 0.1, "gc.precision_theta": torch.rand(nk, 1, generator=g, dtype=dtype) * 0.9 + 0.1}
    return image, Y, idx, alpha, gp


GC_CASES = [(4, 36, 16, 8, 2048), (3, 36, 16, 8, 1024), (2, 51, 19, 4, 512), (2, 51, 36, 32, 1024), (1, 100, 32, 8, 512),
            (3, 12, 5, 4, 32), (2, 12, 5, 4, 16), (2, 7, 7, 2, 24)]


@pytest.mark.parametrize("B,K,nb,nk,out_dim", GC_CASES)
@pytest.mark.parametrize("use_alpha", [True, False])
def test_graphconv_fwd(kn, B, K, nb, nk, out_dim, use_alpha):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=K + nb)
    ref = torch.relu(_gc_reference(Y, idx, alpha if use_alpha else None, image, gp, "gc", nk))
    gauss = _pack_gauss(gp, "gc")
    out = kn.graphconv_fwd(Y.float().view(B * K, -1).to(DEV), idx.int().to(DEV), alpha.float().to(DEV) if use_alpha else None,
                           image.float().to(DEV), gauss, B, K, relu=True)
    assert rel_err(out.view(B, K, -1).cpu(), ref) < 2e-5


def test_graphconv_fused_dropout(kn):
    B, K, nb, nk, out_dim = 6, 36, 16, 8, 2048
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=11)
    args = (Y.float().view(B * K, -1).to(DEV), idx.int().to(DEV), alpha.float().to(DEV), image.float().to(DEV), _pack_gauss(gp, "gc"), B, K)
    base = kn.graphconv_fwd(*args, relu=True)
    dropped = kn.graphconv_fwd(*args, relu=True, dropout_p=0.5, seed=42, offset=3)
    kept = dropped != 0
    assert torch.allclose(dropped[kept], 2.0 * base[kept])
    pos = base > 0
    assert abs((kept & pos).sum().item() / pos.sum().item() - 0.5) < 0.01
    assert torch.equal(dropped, kn.graphconv_fwd(*args, relu=True, dropout_p=0.5, seed=42, offset=3))


@pytest.mark.parametrize("B,K,nb,nk,out_dim", GC_CASES)
def test_graphconv_pool_fwd(kn, B, K, nb, nk, out_dim):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=2 * K + nb)
    g2 = torch.relu(_gc_reference(Y, idx, None, image, gp, "gc", nk))
    pooled_ref, arg_ref = g2.max(1)
    q = torch.randn(B, out_dim, dtype=torch.float64)
    pooled, arg, hq = kn.graphconv_pool_fwd(Y.float().view(B * K, -1).to(DEV), idx.int().to(DEV), image.float().to(DEV),
                                            _pack_gauss(gp, "gc"), q.float().to(DEV), B, K)
    assert rel_err(pooled.cpu(), pooled_ref) < 2e-5
    assert rel_err(hq.cpu(), torch.relu(q) * pooled_ref) < 2e-5
    assert arg.dtype == torch.int64
    top2 = g2.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-4 * top2[:, 0].abs().clamp(min=1e-3)
    assert torch.equal(arg.cpu()[safe], arg_ref[safe])
    # all-zero columns (every node <= 0 after ReLU): first index, as torch.max on CPU
    zero_cols = pooled_ref == 0
    assert (arg.cpu()[zero_cols] == 0).all()


@pytest.mark.parametrize("B,K,nb,nk,out_dim", GC_CASES)
@pytest.mark.parametrize("mode", ["dense_alpha", "dense", "pooled"])
def test_graphconv_bwd(kn, B, K, nb, nk, out_dim, mode):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=3 * K + nb)
    Yr = Y.clone().requires_grad_(True)
    ar = alpha.clone().requires_grad_(True)
    gpr = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    use_alpha = mode == "dense_alpha"
    out = _gc_reference(Yr, idx, ar if use_alpha else None, image, gpr, "gc", nk)
    g = torch.Generator().manual_seed(1)
    gauss = _pack_gauss(gp, "gc")
    common = (Y.float().view(B * K, -1).to(DEV), idx.int().to(DEV), alpha.float().to(DEV) if use_alpha else None, image.float().to(DEV), gauss, B, K)
    if mode == "pooled":
        arg = torch.randint(0, K, (B, out_dim), generator=g)
        dpooled = torch.randn(B, out_dim, generator=g, dtype=torch.float64)
        dO = torch.zeros(B, K, out_dim, dtype=torch.float64).scatter_(1, arg.unsqueeze(1), dpooled.unsqueeze(1))
        dY, dalpha, dgauss = kn.graphconv_bwd(*common, dpooled=dpooled.float().to(DEV), argmax=arg.to(DEV))
    else:
        dO = torch.randn(B, K, out_dim, generator=g, dtype=torch.float64)
        dY, dalpha, dgauss = kn.graphconv_bwd(*common, dO=dO.float().view(B * K, -1).to(DEV))
    wanted = [Yr] + [gpr[f"gc.{k}"] for k in ("mean_rho", "precision_rho", "mean_theta", "precision_theta")] + ([ar] if use_alpha else [])
    grads = torch.autograd.grad((out * dO).sum(), wanted)
    assert rel_err(dY.view(B, K, -1).cpu(), grads[0]) < 2e-5
    dg_ref = torch.cat([x.reshape(-1) for x in grads[1:5]])
    assert rel_err(dgauss.cpu(), dg_ref) < 2e-4
    if use_alpha:
        assert rel_err(dalpha.cpu(), grads[5]) < 2e-5
    else:
        assert dalpha is None


def test_gemm_row_tile_gate(kn):
    """vqa_gemm_bf16s tile_gate: row tiles whose gate value is <= t are not computed (left untouched), the others are exact."""
    M, N, K = 520, 384, 256
    g = torch.Generator().manual_seed(5)
    a = torch.randn(M, K, generator=g).to(DEV); b = torch.randn(N, K, generator=g).to(DEV)
    As, Bs = kn.split(a), kn.split(b)
    full = kn.gemm_s(As, Bs, tile_n=128)
    gate = torch.tensor([9, 4, 7, 4, 2], dtype=torch.int32, device=DEV)          # 5 row tiles of 128
    out = torch.full((M, N), -7.0, device=DEV)
    kn.gemm_s(As, Bs, out=out, tile_n=128, row_gate=(gate, 4))
    for mt, gv in enumerate(gate.tolist()):
        rows = slice(mt * 128, min(M, (mt + 1) * 128))
        if gv <= 4:
            assert torch.all(out[rows] == -7.0), mt
        else:
            assert torch.equal(out[rows], full[rows]), mt
    acc = full.clone()                                                            # accumulate mode (the GRU backward's use)
    kn.gemm_s(As, Bs, out=acc, accumulate=True, split_k=2, tile_n=128, row_gate=(gate, 4))
    assert torch.equal(acc[128:256], full[128:256]) and rel_err(acc[0:128], 2 * full[0:128]) < 1e-6


@pytest.mark.parametrize("B,K,nb,nk,out_dim", [(4, 36, 16, 8, 1024), (2, 51, 19, 4, 512), (2, 100, 32, 8, 1024), (3, 12, 5, 2, 272)])
@pytest.mark.parametrize("with_lo", [True, False])
def test_graphconv_pool_bwd_data(kn, B, K, nb, nk, out_dim, with_lo):
    """Backward data path of the pooled layer as a scatter of coef * dpooled vs the fp64 transposed aggregate."""
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=K + nb)
    Yr = Y.clone().requires_grad_(True)
    out = _gc_reference(Yr, idx, None, image, gp, "gc", nk)
    g = torch.Generator().manual_seed(2)
    arg = torch.randint(0, K, (B, out_dim), generator=g)
    dpooled = torch.randn(B, out_dim, generator=g, dtype=torch.float64)
    dO = torch.zeros(B, K, out_dim, dtype=torch.float64).scatter_(1, arg.unsqueeze(1), dpooled.unsqueeze(1))
    ref, = torch.autograd.grad((out * dO).sum(), [Yr])
    idx_d, img_d = idx.int().to(DEV), image.float().to(DEV)
    ec = kn.graphconv_edge_coef(idx_d, None, img_d, _pack_gauss(gp, "gc"), B, K)
    dY = kn.graphconv_pool_bwd_data_s(dpooled.float().to(DEV), arg.to(DEV), idx_d, ec, B, K, out_dim, with_lo=with_lo)
    assert (dY.lo is None) == (not with_lo)
    assert rel_err(dY.float().view(B, K, -1).cpu(), ref) < (2e-5 if with_lo else 6e-3)


# ---- tensor-core aggregate on split planes (graphconv_mma.cu) -----------------------------------------------------
MMA_CASES = [(4, 36, 16, 8, 2048), (3, 36, 16, 8, 1024), (2, 51, 19, 4, 512), (3, 51, 19, 8, 1024), (2, 51, 19, 8, 2048), (2, 100, 32, 8, 1024), (3, 12, 5, 2, 256), (2, 7, 7, 2, 256)]


@pytest.mark.parametrize("B,K,nb,nk,out_dim", MMA_CASES)
@pytest.mark.parametrize("use_alpha", [True, False])
def test_graphconv_mma_fwd(kn, B, K, nb, nk, out_dim, use_alpha):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=K + nb)
    ref = torch.relu(_gc_reference(Y, idx, alpha if use_alpha else None, image, gp, "gc", nk))
    Ys = kn.split(Y.float().view(B * K, -1).to(DEV))
    out = kn.graphconv_fwd_s(Ys, idx.int().to(DEV), alpha.float().to(DEV) if use_alpha else None, image.float().to(DEV),
                             _pack_gauss(gp, "gc"), B, K, relu=True)
    assert out.shape == (B * K, out_dim)
    assert rel_err(out.float().view(B, K, -1).cpu(), ref) < 3e-5
    # bf16 mode: hi planes only
    out1 = kn.graphconv_fwd_s(kn.split(Y.float().view(B * K, -1).to(DEV), with_lo=False), idx.int().to(DEV),
                              alpha.float().to(DEV) if use_alpha else None, image.float().to(DEV), _pack_gauss(gp, "gc"), B, K)
    assert out1.lo is None and rel_err(out1.float().view(B, K, -1).cpu(), ref) < 2e-2


def test_graphconv_mma_fused_dropout(kn):
    B, K, nb, nk, out_dim = 6, 36, 16, 8, 2048
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=11)
    args = (kn.split(Y.float().view(B * K, -1).to(DEV)), idx.int().to(DEV), alpha.float().to(DEV), image.float().to(DEV), _pack_gauss(gp, "gc"), B, K)
    base = kn.graphconv_fwd_s(*args, relu=True).float()
    dropped = kn.graphconv_fwd_s(*args, relu=True, dropout_p=0.5, seed=42, offset=3).float()
    kept = dropped != 0
    assert torch.allclose(dropped[kept], 2.0 * base[kept], rtol=1e-4, atol=1e-6)
    pos = base > 0
    assert abs((kept & pos).sum().item() / pos.sum().item() - 0.5) < 0.01
    assert torch.equal(dropped, kn.graphconv_fwd_s(*args, relu=True, dropout_p=0.5, seed=42, offset=3).float())
    other = kn.graphconv_fwd_s(*args, relu=True, dropout_p=0.5, seed=42, offset=4).float()
    assert not torch.equal(dropped, other)
    step = torch.tensor(5, dtype=torch.int64, device=DEV)          # device-side step counter shifts the stream
    s5 = kn.graphconv_fwd_s(*args, relu=True, dropout_p=0.5, seed=42, offset=3, step=step).float()
    assert not torch.equal(dropped, s5)
    # masks of neighbouring rows / columns are uncorrelated
    m = kept[:, :-1] & pos[:, :-1] & pos[:, 1:]
    agree = ((kept[:, :-1] == kept[:, 1:]) & pos[:, :-1] & pos[:, 1:]).sum().item() / max(1, (pos[:, :-1] & pos[:, 1:]).sum().item())
    assert abs(agree - 0.5) < 0.01


@pytest.mark.parametrize("B,K,nb,nk,out_dim", MMA_CASES)
def test_graphconv_mma_pool_fwd(kn, B, K, nb, nk, out_dim):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=2 * K + nb)
    g2 = torch.relu(_gc_reference(Y, idx, None, image, gp, "gc", nk))
    pooled_ref, arg_ref = g2.max(1)
    q = torch.randn(B, out_dim, dtype=torch.float64)
    pooled, arg, hq = kn.graphconv_pool_fwd_s(kn.split(Y.float().view(B * K, -1).to(DEV)), idx.int().to(DEV), image.float().to(DEV),
                                              _pack_gauss(gp, "gc"), q.float().to(DEV), B, K)
    assert rel_err(pooled.cpu(), pooled_ref) < 3e-5
    assert rel_err(hq.cpu(), torch.relu(q) * pooled_ref) < 3e-5
    assert arg.dtype == torch.int64
    top2 = g2.topk(2, dim=1).values
    safe = (top2[:, 0] - top2[:, 1]) > 1e-3 * top2[:, 0].abs().clamp(min=1e-3)
    assert torch.equal(arg.cpu()[safe], arg_ref[safe])
    zero_cols = pooled_ref == 0
    assert (arg.cpu()[zero_cols] == 0).all()


@pytest.mark.parametrize("B,K,nb,nk,out_dim", MMA_CASES)
@pytest.mark.parametrize("use_alpha", [True, False])
def test_graphconv_mma_bwd_data(kn, B, K, nb, nk, out_dim, use_alpha):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=3 * K + nb)
    Yr = Y.clone().requires_grad_(True)
    out = _gc_reference(Yr, idx, alpha if use_alpha else None, image, gp, "gc", nk)
    dO = torch.randn(B, K, out_dim, generator=torch.Generator().manual_seed(1), dtype=torch.float64)
    (dY_ref,) = torch.autograd.grad((out * dO).sum(), [Yr])
    dY = kn.graphconv_bwd_data_s(kn.split(dO.float().view(B * K, -1).to(DEV)), idx.int().to(DEV), alpha.float().to(DEV) if use_alpha else None,
                                 image.float().to(DEV), _pack_gauss(gp, "gc"), B, K)
    assert rel_err(dY.float().view(B, K, -1).cpu(), dY_ref) < 3e-5


# the edge kernel stacks G = min(floor(128 / K), 4, B) consecutive images in one MMA: batches that end inside a group, the cap of
# four, one image per group (K > 64), groups whose rows straddle the 32-lane quarters of TMEM, nb not a multiple of 4
EDGE_GROUP_CASES = [(5, 36, 16, 8, 1024), (7, 12, 4, 2, 256), (9, 20, 8, 4, 512), (1, 36, 16, 8, 512), (4, 64, 16, 4, 256), (3, 65, 9, 2, 128), (6, 31, 7, 2, 128)]


@pytest.mark.parametrize("B,K,nb,nk,out_dim", MMA_CASES + EDGE_GROUP_CASES)
@pytest.mark.parametrize("mode", ["dense_alpha", "dense", "pooled"])
def test_graphconv_mma_bwd_edges(kn, B, K, nb, nk, out_dim, mode):
    image, Y, idx, alpha, gp = _gc_inputs(B, K, nb, nk, out_dim, seed=3 * K + nb)
    ar = alpha.clone().requires_grad_(True)
    gpr = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
    use_alpha = mode == "dense_alpha"
    out = _gc_reference(Y, idx, ar if use_alpha else None, image, gpr, "gc", nk)
    g = torch.Generator().manual_seed(1)
    Ys = kn.split(Y.float().view(B * K, -1).to(DEV))
    common = (Ys, idx.int().to(DEV), alpha.float().to(DEV) if use_alpha else None, image.float().to(DEV), _pack_gauss(gp, "gc"), B, K)
    if mode == "pooled":
        arg = torch.randint(0, K, (B, out_dim), generator=g)
        dpooled = torch.randn(B, out_dim, generator=g, dtype=torch.float64)
        dO = torch.zeros(B, K, out_dim, dtype=torch.float64).scatter_(1, arg.unsqueeze(1), dpooled.unsqueeze(1))
        dalpha, dgauss = kn.graphconv_bwd_edges_s(*common, dpooled=dpooled.float().to(DEV), argmax=arg.to(DEV))
    else:
        dO = torch.randn(B, K, out_dim, generator=g, dtype=torch.float64)
        dalpha, dgauss = kn.graphconv_bwd_edges_s(*common, dOs=kn.split(dO.float().view(B * K, -1).to(DEV)))
    wanted = [gpr[f"gc.{k}"] for k in ("mean_rho", "precision_rho", "mean_theta", "precision_theta")] + ([ar] if use_alpha else [])
    grads = torch.autograd.grad((out * dO).sum(), wanted)
    dg_ref = torch.cat([x.reshape(-1) for x in grads[:4]])
    assert rel_err(dgauss.cpu(), dg_ref) < 2e-4
    if use_alpha:
        assert rel_err(dalpha.cpu(), grads[4]) < 3e-5
    else:
        assert dalpha is None


def test_gaussian_weights_match_reference_golden(kn):
    g = load_golden("tiny")
    p = golden_params(g)
    w = kn.gaussian_weights(torch.from_numpy(g["layer.nbr_pseudo"]).to(DEV), _pack_gauss(p, "graph_convolution_1"))
    assert rel_err(w.cpu(), g["layer.gauss_w"]) < 1e-5


def test_graphconv_argument_errors(kn):
    image, Y, idx, alpha, gp = _gc_inputs(2, 12, 5, 4, 32)
    gauss = _pack_gauss(gp, "gc")
    with pytest.raises(RuntimeError):   # out_dim / nk not a multiple of 4
        kn.graphconv_fwd(torch.randn(24, 24, device=DEV), idx.int().to(DEV), None, image.float().to(DEV), gauss, 2, 12)
    with pytest.raises(RuntimeError):   # nb > K
        kn.topk_softmax(torch.randn(1, 4, 4, device=DEV), 5)


def test_out_of_range_token_is_flagged_like_nn_embedding(kn):
    """nn.Embedding raises on a token id outside the table (reference sparse_graph_model.py:117); the gather kernel never reads out of
    range, sets a device error bit instead, and the host raises IndexError at its next sync point (kernels.check_device_errors)."""
    wemb = torch.randn(50, 24, device=DEV)
    q = torch.randint(1, 50, (6, 9), device=DEV)
    kn.embed_gather_split(q, wemb, 7)
    kn.check_device_errors()                              # clean
    q[2, 3] = 50
    out = kn.embed_gather_split(q, wemb, 7)
    assert torch.isfinite(out.float()).all()
    with pytest.raises(IndexError):
        kn.check_device_errors()
    kn.check_device_errors()                              # the flag was cleared by the raise
    q[2, 3] = -1
    kn.embed_gather_split(q, wemb, 7)
    with pytest.raises(IndexError):
        kn.check_device_errors()


@pytest.mark.parametrize("n,nb,nk,F", [(72, 16, 8, 2052), (24, 5, 4, 20), (9, 19, 32, 132), (5, 3, 2, 7)])
def test_patch_operator_fwd_bwd(kn, n, nb, nk, F):
    """The layer API's patch operator (reference layers.py:136-137: torch.bmm(weights^T, neighbourhood)) and its backward vs fp64."""
    g = torch.Generator().manual_seed(n + nb)
    X = torch.randn(n, nb, F, generator=g, dtype=torch.float64, requires_grad=True)
    w = torch.rand(n, nb, nk, generator=g, dtype=torch.float64, requires_grad=True)
    Z = torch.bmm(w.transpose(1, 2), X)
    dZ = torch.randn(n, nk, F, generator=g, dtype=torch.float64)
    gX, gw = torch.autograd.grad((Z * dZ).sum(), [X, w])
    Zc = kn.patch_operator_fwd(X.detach().float().to(DEV), w.detach().float().to(DEV))
    assert rel_err(Zc.cpu(), Z.detach()) < 2e-6
    dX, dw = kn.patch_operator_bwd(X.detach().float().to(DEV), w.detach().float().to(DEV), dZ.float().to(DEV))
    assert rel_err(dX.cpu(), gX) < 2e-6 and rel_err(dw.cpu(), gw) < 5e-6
