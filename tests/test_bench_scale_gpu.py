"""Parity AT BENCH SCALE: the kernels `bench.py` actually runs, against fp64 restatements on the same inputs.

The small-shape suites (test_kernels_gpu.py, test_model_gpu.py) never reach the code paths a B=512 step takes:
`gemm_bf16s_persistent_kernel` needs more tiles than SMs (> 74 CTA pairs / > 148 CTAs), `agg_persistent_kernel` only
wraps its rings and double buffers when a CTA walks several (image, kernel-slab) items (B * slabs > 148), the split-K
weight-gradient products and the two-wave edge kernel only exist at B*K ~ 18k rows.  Every test here is sized so that
those paths are the ones that run (asserted where the dispatch rule is simple), and compared with

  * fp64 torch expressions evaluated on the GPU for the dense products,
  * the oracle's graph-convolution restatement (oracle/vqa_oracle.py, pinned to the reference by tests/golden) in fp64,
    chunked over images, for the four graph kernels,
  * the oracle's full train step (forward + loss + every parameter gradient) in fp64 at the FULL widths of BASELINE.json
    configs[1] (VQA2: B=512, K=36, F=2052, hid 1024, 3000 answers), configs[3] (medical: K=51, F=1028, nb=19) and
    configs[4] (K=100, nb=32).

The checker runs on the GPU only because fp64 at these sizes takes minutes on the host; it is stock torch, none of this
repo's kernels.  Tolerance: 1e-3 max-norm relative (BASELINE.json north_star, fp32 mode); measured errors are printed.
"""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_err
from oracle import vqa_oracle as O
from vqa_b200.synthetic import WORKLOADS, Workload, make_batch, make_wemb

pytestmark = pytest.mark.gpu
DEV = "cuda"
TOL = 1e-3


@pytest.fixture(scope="module")
def kn():
    from vqa_b200 import kernels
    return kernels


def _rel(a, b):
    """max-norm relative error, evaluated on the device (tensors here have up to 4e7 elements)."""
    b = b.double()
    den = b.abs().max().item()
    return (a.double() - b).abs().max().item() / (den if den > 0 else 1.0)


def _randn(shape, seed, scale=1.0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return torch.randn(shape, generator=g, device=DEV) * scale


# --------------------------------------------------------------------------------------------- dense products
# name, M, N, K, a_mn, b_mn, tile_n, split_k, epilogue; CTA pairs when tile 256 and >= 8 tile rows
BENCH_GEMMS = [
    ("Y1 = X Wc1^T (planes out)", 18432, 2048, 2052, False, False, 0, 1, "planes"),
    ("Y2 = G1 Wc2^T (planes out)", 18432, 1024, 2048, False, False, 0, 1, "planes"),
    ("dG1 = dY2 Wc2 masked by G1 (planes out)", 18432, 2048, 1024, False, True, 0, 1, "mask_planes"),
    ("h1 = relu(X W1x^T + b + q-term) (planes out)", 18432, 512, 2052, False, False, 0, 1, "gl1"),
    ("h1, 128-wide tiles", 18432, 512, 2052, False, False, 128, 1, "gl1"),
    ("h2 = relu(h1 W2^T + b)", 18432, 512, 512, False, False, 0, 1, "bias_relu"),
    ("dWc1 = dY1^T X", 2048, 2052, 18432, True, True, 0, 1, "plain"),
    ("dWc1, split-K 3 (persistent + atomics)", 2048, 2052, 18432, True, True, 0, 3, "plain"),
    ("dWc2 = dY2^T G1", 1024, 2048, 18432, True, True, 0, 2, "plain"),
    ("dW1x = dh1^T X", 512, 2052, 18432, True, True, 0, 4, "plain"),
    ("GI = E W_ih^T + b (all steps)", 7168, 3072, 300, False, False, 0, 1, "bias"),
    ("MED Y1", 26112, 2048, 1028, False, False, 0, 1, "planes"),
    ("MED dWc1", 2048, 1028, 26112, True, True, 0, 2, "plain"),
]


def _is_persistent(M, N, K, tile_n, split_k):
    """csrc/gemm_bf16s.cu dispatch, restated: persistent when there are more (pairs of) tiles than SMs (pairs)."""
    mt = (M + 127) // 128
    bn = tile_n
    if bn == 0:
        ctas = lambda wdt: mt * ((N + wdt - 1) // wdt) * split_k
        bn = 256 if N > 128 else (128 if N > 64 else 64)
        if bn == 256 and ctas(256) < 120:
            bn = 128
        if bn == 128 and ctas(128) < 120:
            bn = 64
    cl = 2 if (bn == 256 and mt >= 8) else 1
    if cl == 2:
        mt = (mt + 1) & ~1
    units = ((N + bn - 1) // bn) * (mt // cl) * split_k
    return bn >= 128 and units > 148 // cl


@pytest.mark.parametrize("name,M,N,K,a_mn,b_mn,tile_n,split_k,epi", BENCH_GEMMS, ids=[c[0] for c in BENCH_GEMMS])
@pytest.mark.parametrize("passes,tol", [(3, 3e-5), (1, 8e-3)])
def test_gemm_at_bench_shapes_matches_fp64(kn, name, M, N, K, a_mn, b_mn, tile_n, split_k, epi, passes, tol):
    a = _randn((K, M) if a_mn else (M, K), M + 3 * N + K)
    b = _randn((K, N) if b_mn else (N, K), 7 * M + N + K)
    ad, bd = a.double(), b.double()
    ref = (ad.t() if a_mn else ad) @ (bd if b_mn else bd.t())
    As, Bs = kn.split(a, with_lo=passes == 3), kn.split(b, with_lo=passes == 3)
    kw = dict(a_mn=a_mn, b_mn=b_mn, passes=passes, tile_n=tile_n, split_k=split_k)
    if epi in ("planes", "mask_planes", "gl1"):
        out_s = kn.empty_split(M, N, DEV, with_lo=passes == 3)
        if epi == "mask_planes":
            mask = _randn((M, N), 5)
            ref = torch.where(mask.bfloat16().float() > 0, 2.0 * ref, torch.zeros((), dtype=torch.float64, device=DEV))
            kn.gemm_s(As, Bs, out_split=out_s, want_f32=False, aux=kn.split(mask, with_lo=passes == 3), aux_scale=2.0, **kw)
        elif epi == "gl1":
            bias, rb = _randn((N,), 6), _randn((M // 36, N), 7)
            ref = (ref + bias.double() + rb.double().repeat_interleave(36, 0)).clamp_(min=0)
            kn.gemm_s(As, Bs, out_split=out_s, want_f32=False, bias=bias, rowbcast=rb, group=36, relu=True, **kw)
        else:
            kn.gemm_s(As, Bs, out_split=out_s, want_f32=False, **kw)
        got = out_s.float()
    elif epi in ("bias", "bias_relu"):
        bias = _randn((N,), 6)
        ref = ref + bias.double()
        if epi == "bias_relu":
            ref.clamp_(min=0)
        got = kn.gemm_s(As, Bs, bias=bias, relu=epi == "bias_relu", **kw)
    else:
        got = kn.gemm_s(As, Bs, **kw)
    torch.cuda.synchronize()
    err = _rel(got, ref)
    if passes == 3 and K > 8192:
        # one fp32 TMEM accumulator over K = 18432 .. 26112: tcgen05 accumulates with truncation, ~2^-24 of the running sum per
        # 16-deep MMA step (K/16 steps) -> measured 3.5e-5 .. 7e-5 here; the split-K variants are proportionally closer
        tol = 1.5e-4
    print(f"{name}: {'persistent' if _is_persistent(M, N, K, tile_n, split_k) else 'one tile per CTA'}, passes={passes}, rel err {err:.2e}")
    # bf16 planes of the OUTPUT carry 16 (3-pass) / 8 (1-pass) mantissa bits: part of the stated budget
    assert err < tol * (2.0 if passes == 1 and epi in ("planes", "mask_planes", "gl1") else 1.0)


def test_bench_gemm_cases_reach_both_persistent_instantiations():
    pers = [c for c in BENCH_GEMMS if _is_persistent(*c[1:4], c[6], c[7])]
    assert any(c[6] == 128 for c in pers) and any(c[6] == 0 and c[2] >= 1024 for c in pers) and any(c[7] > 1 for c in pers)
    assert len(pers) >= 8


# --------------------------------------------------------------------------------------------- graph kernels
def _gc_inputs_dev(B, K, nb, nk, out_dim, seed):
    g = torch.Generator(device=DEV).manual_seed(seed)
    f64 = dict(generator=g, device=DEV, dtype=torch.float64)
    image = torch.rand(B, K, 12, **f64)
    xy1 = torch.rand(B, K, 2, **f64) * 0.7
    image[..., -4:-2] = xy1
    image[..., -2:] = xy1 + torch.rand(B, K, 2, **f64) * 0.25 + 0.05
    Y = torch.randn(B, K, out_dim, **f64)
    idx = torch.rand(B, K, K, generator=g, device=DEV).argsort(-1)[..., :nb].contiguous()      # nb distinct neighbours per node
    alpha = torch.softmax(torch.randn(B, K, nb, **f64), -1)
    gp = {"gc.mean_rho": torch.rand(nk, 1, **f64), "gc.mean_theta": (torch.rand(nk, 1, **f64) * 2 - 1) * math.pi,
          "gc.precision_rho": torch.rand(nk, 1, **f64) * 0.9 + 0.1, "gc.precision_theta": torch.rand(nk, 1, **f64) * 0.9 + 0.1}
    return image, Y, idx, alpha, gp


def _pack(gp):
    return torch.cat([gp[f"gc.{k}"].reshape(-1) for k in ("mean_rho", "precision_rho", "mean_theta", "precision_theta")]).float().contiguous()


def _gc_ref(Y, idx, alpha, image, gp, nk):
    """The oracle's aggregate on fp64 device tensors (same statements as test_kernels_gpu._gc_reference)."""
    B, K, nb = idx.shape
    pseudo = O.gather_pseudo(O.polar_pseudo_coordinates(O.box_centres(image)), idx)
    w = O.gaussian_kernel_weights(pseudo, gp, "gc").view(B, K, nb, nk)
    D = Y.shape[-1] // nk
    nbr = O.gather_neighbours(Y, idx).view(B, K, nb, nk, D)
    coef = w if alpha is None else w * alpha.unsqueeze(-1)
    return (coef.unsqueeze(-1) * nbr).sum(2).reshape(B, K, nk * D)


def _chunks(B, n=32):
    return [slice(i, min(B, i + n)) for i in range(0, B, n)]


# B is chosen so that B * slabs > 2 * 148: every persistent CTA walks >= 2 items (ring wrap, coefficient double buffer)
SCALE_CASES = [(160, 36, 16, 8, 2048), (320, 36, 16, 8, 1024), (112, 51, 19, 8, 2048), (224, 51, 19, 8, 1024), (48, 100, 32, 8, 2048)]


def _items_per_cta(B, nk, out_dim):
    tpk = (out_dim // nk) // 128
    nkc = min(nk, 4, (4 + tpk - 1) // tpk)
    while nk % nkc:
        nkc -= 1
    return B * (nk // nkc) / 148.0


@pytest.mark.parametrize("B,K,nb,nk,out_dim", SCALE_CASES)
def test_graphconv_fwd_multi_item(kn, B, K, nb, nk, out_dim):
    assert _items_per_cta(B, nk, out_dim) > (2.0 if K <= 64 else 1.0)
    image, Y, idx, alpha, gp = _gc_inputs_dev(B, K, nb, nk, out_dim, seed=K + nb)
    Ys = kn.split(Y.float().view(B * K, -1))
    args = (idx.int(), alpha.float(), image.float(), _pack(gp), B, K)
    out = kn.graphconv_fwd_s(Ys, *args, relu=True).float().view(B, K, -1)
    worst = 0.0
    for c in _chunks(B):
        ref = torch.relu(_gc_ref(Y[c], idx[c], alpha[c], image[c], gp, nk))
        worst = max(worst, _rel(out[c], ref))
    print(f"fwd B={B} K={K} out={out_dim}: {_items_per_cta(B, nk, out_dim):.1f} items per CTA, rel err {worst:.2e}")
    assert worst < 3e-5
    # fused dropout at the same scale: kept entries are exactly the un-dropped ones times 1/(1-p), half of the positives survive
    dropped = kn.graphconv_fwd_s(Ys, *args, relu=True, dropout_p=0.5, seed=42, offset=3).float().view(B, K, -1)
    kept, pos = dropped != 0, out > 0
    assert torch.allclose(dropped[kept], 2.0 * out[kept], rtol=1e-4, atol=1e-6)
    assert abs((kept & pos).sum().item() / pos.sum().item() - 0.5) < 2e-3
    assert torch.equal(dropped, kn.graphconv_fwd_s(Ys, *args, relu=True, dropout_p=0.5, seed=42, offset=3).float().view(B, K, -1))


@pytest.mark.parametrize("B,K,nb,nk,out_dim", SCALE_CASES)
def test_graphconv_pool_fwd_multi_item(kn, B, K, nb, nk, out_dim):
    image, Y, idx, alpha, gp = _gc_inputs_dev(B, K, nb, nk, out_dim, seed=2 * K + nb)
    q = torch.randn(B, out_dim, device=DEV, dtype=torch.float64, generator=torch.Generator(device=DEV).manual_seed(4))
    pooled, arg, hq = kn.graphconv_pool_fwd_s(kn.split(Y.float().view(B * K, -1)), idx.int(), image.float(), _pack(gp), q.float(), B, K)
    worst, flips, n = 0.0, 0, 0
    for c in _chunks(B):
        g2 = torch.relu(_gc_ref(Y[c], idx[c], None, image[c], gp, nk))
        p_ref, a_ref = g2.max(1)
        worst = max(worst, _rel(pooled[c], p_ref), _rel(hq[c], torch.relu(q[c]) * p_ref))
        top2 = g2.topk(2, dim=1).values
        safe = (top2[:, 0] - top2[:, 1]) > 1e-3 * top2[:, 0].abs().clamp(min=1e-3)
        flips += int((arg[c][safe] != a_ref[safe]).sum())
        n += int(safe.sum())
        # columns in which every node is <= 0 after the ReLU: exact zeros on both sides -> first index (torch.max); a column the fp32
        # evaluation makes barely positive (cancellation noise around 0) is a near-tie like any other
        zero = (p_ref == 0) & (pooled[c] == 0)
        assert (arg[c][zero] == 0).all()
        assert pooled[c][p_ref == 0].abs().max().item() <= 1e-5 * p_ref.max().item() if (p_ref == 0).any() else True
    print(f"pool fwd B={B} K={K} out={out_dim}: rel err {worst:.2e}, arg-max mismatches on {n} safe columns: {flips}")
    assert worst < 3e-5 and flips == 0


@pytest.mark.parametrize("B,K,nb,nk,out_dim", SCALE_CASES)
def test_graphconv_bwd_multi_item(kn, B, K, nb, nk, out_dim):
    """dY (bwd_data), dalpha and the Gaussian-parameter gradients (bwd_edges) of the dense layer, and both for the pooled layer."""
    image, Y, idx, alpha, gp = _gc_inputs_dev(B, K, nb, nk, out_dim, seed=3 * K + nb)
    g = torch.Generator(device=DEV).manual_seed(1)
    dO = torch.randn(B, K, out_dim, generator=g, device=DEV, dtype=torch.float64)
    arg = torch.randint(0, K, (B, out_dim), generator=g, device=DEV)
    dpooled = torch.randn(B, out_dim, generator=g, device=DEV, dtype=torch.float64)
    dO_pool = torch.zeros(B, K, out_dim, dtype=torch.float64, device=DEV).scatter_(1, arg.unsqueeze(1), dpooled.unsqueeze(1))
    names = ("mean_rho", "precision_rho", "mean_theta", "precision_theta")
    idx32, img32, gauss = idx.int(), image.float(), _pack(gp)
    Ys = kn.split(Y.float().view(B * K, -1))
    # ---- the CUDA path
    dY = kn.graphconv_bwd_data_s(kn.split(dO.float().view(B * K, -1)), idx32, alpha.float(), img32, gauss, B, K).float().view(B, K, -1)
    dalpha, dgauss = kn.graphconv_bwd_edges_s(Ys, idx32, alpha.float(), img32, gauss, B, K, dOs=kn.split(dO.float().view(B * K, -1)))
    ec = kn.graphconv_edge_coef(idx32, None, img32, gauss, B, K)
    dY_pool = kn.graphconv_pool_bwd_data_s(dpooled.float(), arg, idx32, ec, B, K, out_dim).float().view(B, K, -1)
    _, dgauss_pool = kn.graphconv_bwd_edges_s(Ys, idx32, None, img32, gauss, B, K, dpooled=dpooled.float(), argmax=arg)
    # ---- fp64 autograd through the oracle's statements, chunked over images
    dg_ref = torch.zeros(4 * nk, dtype=torch.float64, device=DEV)
    dg_pool_ref = torch.zeros_like(dg_ref)
    e_dy = e_da = e_dyp = 0.0
    for c in _chunks(B):
        Yr, ar = Y[c].clone().requires_grad_(True), alpha[c].clone().requires_grad_(True)
        gpr = {k: v.clone().requires_grad_(True) for k, v in gp.items()}
        out = _gc_ref(Yr, idx[c], ar, image[c], gpr, nk)
        grads = torch.autograd.grad((out * dO[c]).sum(), [Yr, ar] + [gpr[f"gc.{k}"] for k in names])
        e_dy, e_da = max(e_dy, _rel(dY[c], grads[0])), max(e_da, _rel(dalpha[c], grads[1]))
        dg_ref += torch.cat([x.reshape(-1) for x in grads[2:]])
        out_p = _gc_ref(Yr, idx[c], None, image[c], gpr, nk)
        grads_p = torch.autograd.grad((out_p * dO_pool[c]).sum(), [Yr] + [gpr[f"gc.{k}"] for k in names])
        e_dyp = max(e_dyp, _rel(dY_pool[c], grads_p[0]))
        dg_pool_ref += torch.cat([x.reshape(-1) for x in grads_p[1:]])
    e_g, e_gp = _rel(dgauss, dg_ref), _rel(dgauss_pool, dg_pool_ref)
    print(f"bwd B={B} K={K} out={out_dim}: dY {e_dy:.2e} dalpha {e_da:.2e} dgauss {e_g:.2e} | pooled: dY {e_dyp:.2e} dgauss {e_gp:.2e}")
    assert e_dy < 3e-5 and e_da < 3e-5 and e_dyp < 3e-5
    assert e_g < 2e-4 and e_gp < 2e-4


# --------------------------------------------------------------------------------------------- top-k on reference adjacencies
@pytest.mark.parametrize("name,nb", [("topk_k51", 19), ("topk_k100", 32)])
def test_topk_indices_bit_exact_given_reference_adjacency_k51_k100(kn, name, nb):
    """north_star: neighbour indices bit-exact GIVEN the reference's adjacency, at the node counts of configs[3] / configs[4]
    (fixtures: adjacency + neighbour sets produced by the unmodified reference, tests/golden/make_golden.py::topk_fixtures)."""
    g = load_golden(name)
    adj = torch.from_numpy(g["out.adjacency"]).to(DEV)
    assert g["nbr.idx_sorted"].shape[-1] == nb
    idx, alpha = kn.topk_softmax(adj, nb)
    order = idx.long().argsort(-1)
    assert np.array_equal(torch.gather(idx.long(), -1, order).cpu().numpy(), g["nbr.idx_sorted"])
    assert rel_err(torch.gather(alpha, -1, order).cpu(), g["nbr.alpha_sorted"]) < 1e-6
    vals = torch.gather(adj, -1, idx.long())
    assert (vals[..., :-1] >= vals[..., 1:]).all()


# --------------------------------------------------------------------------------------------- the whole step
def _oracle_train_step_fp64(params, batch, w, chunk=None):
    """oracle.train_step_grads in fp64 on the device, chunked over images (the loss is a mean over the batch, so the chunk
    gradients add with weight n_chunk / B).  Returns loss, grads, logits, adjacency, arg-max top-2 gaps."""
    p64 = {k: v.double().to(DEV) for k, v in params.items()}
    q, img, tgt = batch["question"].to(DEV), batch["image"].double().to(DEV), batch["target"].double().to(DEV)
    qlen = [int(x) for x in batch["qlen"]]
    B = q.shape[0]
    chunk = chunk or (64 if w.n_obj <= 64 else 32)               # the reference materialises (B, K, nb, F) neighbourhoods: 3.4 GB per 32 images at K=100
    grads = {k: torch.zeros_like(v) for k, v in p64.items()}
    loss, logits, adj = 0.0, [], []
    for c in _chunks(B, chunk):
        n = c.stop - c.start
        l, g, (lg, a, _) = O.train_step_grads(p64, q[c], img[c], qlen[c], tgt[c], w.neighbourhood, w.n_kernels)
        loss += float(l) * n / B
        for k in grads:
            grads[k] += g[k] * (n / B)
        logits.append(lg); adj.append(a)
    return loss, grads, torch.cat(logits), torch.cat(adj)


def _safe_batch(model, w, params, want, seed):
    """`want` images on which the CUDA path and the fp64 oracle select the SAME neighbourhoods.  A node whose nb-th and (nb+1)-th
    adjacency entries differ by less than the evaluation error may legitimately end up with either neighbour (ties /
    near-ties are any valid choice; on this synthetic data 1 % of the rows have a relative margin below 4e-5, and the reference's own
    fp32 run differs from fp64 on some of them too); such an image then computes a different - equally valid - function, so it is
    taken out of the comparison: ~15 % more images are generated, the forward of the CUDA path and the oracle's graph learner run on
    all of them, and the first `want` images with identical neighbour sets are kept.  Guard against hiding a real defect: every
    dropped image must have a row whose margin is below 1e-4 of its largest adjacency entry, the adjacency itself must agree
    to 1e-4 on ALL images, and at least 75 % must be kept."""
    from vqa_b200 import kernels as kn
    nb = w.neighbourhood
    b = make_batch(w, seed=seed, batch=int(want * 1.3) + 8)
    p64 = {k: v.double().to(DEV) for k, v in params.items()}
    q, img = b["question"].to(DEV), b["image"].to(DEV)
    qlen = [int(x) for x in b["qlen"]]
    with torch.no_grad():
        _, adj, _ = model(q, img, b["K"].to(DEV), b["qlen"])
        ours = kn.topk_softmax(adj, nb)[0].long().sort(-1).values
        same, tight, worst = [], [], 0.0
        for c in _chunks(q.shape[0], 64):
            qenc = O.gru_last_hidden(p64["wembed.weight"][q[c]], qlen[c], p64)
            nodes = torch.cat((img[c].double(), qenc.unsqueeze(1).expand(-1, w.n_obj, -1)), dim=-1)
            a = O.graph_learner(nodes, p64)
            worst = max(worst, _rel(adj[c], a))
            srt = a.sort(dim=-1, descending=True)
            margin = (srt.values[..., nb - 1] - srt.values[..., nb]) / a.abs().amax(dim=(1, 2)).view(-1, 1)
            tight.append(margin.amin(dim=1) < 1e-4)
            same.append((srt.indices[..., :nb].sort(-1).values == ours[c]).all(-1).all(-1))
    same, tight = torch.cat(same).cpu(), torch.cat(tight).cpu()
    assert worst < 1e-4, f"adjacency differs from the oracle by {worst:.2e}"
    assert bool(tight[~same].all()), "an image selected other neighbours although none of its rows is a near-tie"
    keep = same.nonzero().squeeze(1)[:want]
    assert keep.numel() == want and same.float().mean() > 0.75, f"only {int(same.sum())} of {same.numel()} images keep the oracle's neighbour sets"
    dropped = int((~same[:int(keep[-1]) + 1]).sum())
    out = {k: (v[keep] if torch.is_tensor(v) else [v[i] for i in keep.tolist()]) for k, v in b.items()}
    return out, dropped


FULL_WIDTH = {
    # BASELINE.json configs[1]: the bench workload itself
    "vqa2_b512": (WORKLOADS["vqa2_b512"], 512, True),
    # configs[3]: medical feature shapes (K=51, F=1028, nb=19, 512 answers)
    "med_b512": (WORKLOADS["med_b512"], 512, True),
    # configs[4]: K=100, nb=32 (forward is what the config measures; backward checked too, on a smaller batch)
    "eval_k100": (WORKLOADS["eval_k100"], 192, True),
    # medical grid-search corner (run_imageclef.py:218-219): nk=16 -> layer 2 has (out/nk) = 64 and takes the CUDA-core aggregate
    "med_nk16": (Workload("med_nk16", 64, 51, 1028, out_dim=512, neighbourhood=16, n_kernels=16, max_qlen=15, dropout=0.0), 64, True),
    "med_nk4_nb36": (Workload("med_nk4_nb36", 64, 51, 1028, out_dim=512, neighbourhood=36, n_kernels=4, max_qlen=15, dropout=0.0), 64, True),
}


def _forced_oracle_grads(params, batch, w, dec, chunk=None):
    """The oracle's train step in fp64 with every DISCRETE decision taken from the CUDA path's own forward (`dec`): which
    neighbours (top-k), which units are active (the ReLU masks of h1, h2, GC1, the gate and out_1) and which node wins each
    max-pool column.  Identical to oracle.train_step_grads wherever the two forwards decide alike; where a pre-activation lies
    within the forward error (~1e-5) of a kink, either decision is a valid sub-gradient, and taking the same one on both sides
    makes the comparison a test of the backward ARITHMETIC instead of a lottery over mask flips."""
    p64 = {k: v.double().to(DEV) for k, v in params.items()}
    q, img, tgt = batch["question"].to(DEV), batch["image"].double().to(DEV), batch["target"].double().to(DEV)
    qlen = [int(x) for x in batch["qlen"]]
    B, K, nk = q.shape[0], w.n_obj, w.n_kernels
    chunk = chunk or (64 if K <= 64 else 32)
    grads = {k: torch.zeros_like(v) for k, v in p64.items()}
    for c in _chunks(B, chunk):
        p = {k: v.detach().clone().requires_grad_(True) for k, v in p64.items()}
        x, d = img[c], {k: v[c] for k, v in dec.items()}
        pseudo = O.polar_pseudo_coordinates(O.box_centres(x))
        qenc = O.gru_last_hidden(p["wembed.weight"][q[c]], qlen[c], p)
        nodes = torch.cat((x, qenc.unsqueeze(1).expand(-1, K, -1)), dim=-1)
        h = O.wn_linear(nodes, p, "adjacency_1.edge_layer_1") * d["h1"]
        h = O.wn_linear(h, p, "adjacency_1.edge_layer_2") * d["h2"]
        adj = h @ h.transpose(1, 2)
        alpha = torch.softmax(torch.gather(adj, -1, d["idx"]), dim=-1)
        nbp = O.gather_pseudo(pseudo, d["idx"])
        g1 = O.graph_convolution(alpha.unsqueeze(-1) * O.gather_neighbours(x, d["idx"]), nbp, p, "graph_convolution_1", nk) * d["g1"]
        g2 = O.graph_convolution(O.gather_neighbours(g1, d["idx"]), nbp, p, "graph_convolution_2", nk)
        pooled = torch.gather(g2, 1, d["arg"].unsqueeze(1)).squeeze(1) * d["pool"]
        hid = O.wn_linear(qenc * d["q"] * pooled, p, "out_1") * d["o1"]
        loss = O.multilabel_soft_margin_loss(O.wn_linear(hid, p, "out_2"), tgt[c]) * ((c.stop - c.start) / B)
        names = list(p)
        for n, g in zip(names, torch.autograd.grad(loss, [p[n] for n in names], allow_unused=True)):
            if g is not None:
                grads[n] += g
    return grads


@pytest.mark.parametrize("name", list(FULL_WIDTH))
def test_full_width_train_step_matches_oracle(name, monkeypatch):
    """Model.forward + MultiLabelSoftMarginLoss + backward at the full widths of the BASELINE configs against the oracle in fp64
    (reference operation order), dropout off.  Forward: logits, adjacency, loss within 1e-3 (measured ~1e-5).  Backward: every
    parameter gradient within 1e-3 (max-norm relative) of the oracle evaluated with the SAME discrete decisions
    (_forced_oracle_grads).  The comparison with the oracle's own decisions is printed too and only loosely bounded: a ReLU unit whose
    pre-activation is within the forward error of zero switches a whole per-sample gradient path on or off (one flipped out_1 unit
    changes that question's d(loss)/d(q) by ~1/sqrt(1500) = 2.6 %, and rows of wembed.grad ARE per-question gradients), so at
    B = 512 - 1.5 M out_1 units, ~50 of them within 1e-5 of zero - the max-norm over such tensors measures mask flips, not
    arithmetic; the reference's own fp32 run flips ~25x less often (its forward error is 4e-7) but not never."""
    import sparse_graph_model as M
    from vqa_b200 import kernels as kn, ops
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    w, B, with_bwd = FULL_WIDTH[name]
    torch.manual_seed(1000)
    model = M.Model(pretrained_wemb=make_wemb(w), **(w.model_kwargs() | {"dropout": 0.0}))
    with torch.no_grad():   # no degenerate Gaussian widths (0/0 rows are legal in the reference but make every comparison NaN)
        for gc in (model.graph_convolution_1, model.graph_convolution_2):
            gc.precision_rho.clamp_(min=0.05); gc.precision_theta.clamp_(min=0.05)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).train()
    batch, dropped = _safe_batch(model, w, params, B, seed=11)
    q, img, K, tgt = batch["question"].to(DEV), batch["image"].to(DEV), batch["K"].to(DEV), batch["target"].to(DEV)
    seen = {}
    fwd = ops.ConditionedGraphFn.forward

    def recording_forward(ctx, *args):
        out = fwd(ctx, *args)
        seen["ctx"] = ctx
        return out
    monkeypatch.setattr(ops.ConditionedGraphFn, "forward", staticmethod(recording_forward))
    kn_before = kn.LAUNCHES
    logits, adj, arg = model(q, img, K, batch["qlen"])
    loss = MultiLabelSoftMarginLoss()(logits, tgt)
    ctx = seen["ctx"]
    (_, qenc, *_rest, h2, idx, _alpha, pooled, argmax) = ctx.saved_tensors
    Xs, qs, h1s, G1s, hqs, o1s = ctx.splits[:6]
    Kn = w.n_obj
    dec = {"idx": idx.long(), "h1": (h1s.float() > 0).view(B, Kn, -1).double(), "h2": (h2 > 0).view(B, Kn, -1).double(),
           "g1": (G1s.float() > 0).view(B, Kn, -1).double(), "arg": argmax, "pool": (pooled > 0).double(), "q": (qenc > 0).double(),
           "o1": (o1s.float() > 0).double()}
    loss.backward()
    torch.cuda.synchronize()
    assert kn.LAUNCHES > kn_before
    ref_loss, ref_grads, ref_logits, ref_adj = _oracle_train_step_fp64(params, batch, w)
    forced = _forced_oracle_grads(params, batch, w, dec)
    e_log, e_adj = _rel(logits.detach(), ref_logits), _rel(adj.detach(), ref_adj)
    errs = {k: _rel(v.grad, forced[k]) for k, v in model.named_parameters()}
    free = {k: _rel(v.grad, ref_grads[k]) for k, v in model.named_parameters()}
    worst, wfree = max(errs, key=errs.get), max(free, key=free.get)
    short = lambda k: k.replace("graph_convolution", "gc").replace("adjacency_1.edge_layer", "gl").replace("conv_weights.", "w").replace(".weight", ".w")
    print(f"{name}: B={B} ({dropped} images whose neighbour sets hinge on a near-tie replaced), loss {loss.item():.7f} vs {ref_loss:.7f}, "
          f"logits {e_log:.2e}, adjacency {e_adj:.2e}; gradients, same decisions: worst {errs[worst]:.2e} ({worst}); "
          f"oracle's own decisions: worst {free[wfree]:.2e} ({wfree})")
    print("   same decisions: " + ", ".join(f"{short(k)} {e:.1e}" for k, e in errs.items()))
    print("   own decisions:  " + ", ".join(f"{short(k)} {e:.1e}" for k, e in free.items()))
    assert abs(loss.item() - ref_loss) < 1e-5 * abs(ref_loss)
    assert e_log < TOL and e_adj < TOL
    for k, e in errs.items():
        assert e < TOL, (k, e)
    for k, e in free.items():                      # mask flips re-route per-sample gradient paths (see the docstring): bounded, not 1e-3
        assert e < 1e-1, (k, e)
    assert arg.dtype == torch.int64 and arg.shape == (B, w.hid_dim) and int(arg.min()) >= 0 and int(arg.max()) < w.n_obj


def test_full_width_bf16_mode_stated_tolerance():
    """--precision bf16 at the bench workload: 1-pass products outside the graph-learner chain.  Stated tolerance (max-norm
    relative): logits 2e-2, weight gradients 1e-1, Gaussian-parameter gradients 2e-1."""
    import sparse_graph_model as M
    from vqa_b200 import ops
    from vqa_b200.loss import MultiLabelSoftMarginLoss
    w, B = WORKLOADS["vqa2_b512"], 256
    torch.manual_seed(1000)
    model = M.Model(pretrained_wemb=make_wemb(w), **(w.model_kwargs() | {"dropout": 0.0}))
    with torch.no_grad():
        for gc in (model.graph_convolution_1, model.graph_convolution_2):
            gc.precision_rho.clamp_(min=0.05); gc.precision_theta.clamp_(min=0.05)
    params = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.to(DEV).train()
    batch, _ = _safe_batch(model, w, params, B, seed=12)
    ops.set_precision("bf16")
    try:
        logits, adj, _ = model(batch["question"].to(DEV), batch["image"].to(DEV), batch["K"].to(DEV), batch["qlen"])
        MultiLabelSoftMarginLoss()(logits, batch["target"].to(DEV)).backward()
    finally:
        ops.set_precision("fp32")
    _, ref_grads, ref_logits, ref_adj = _oracle_train_step_fp64(params, batch, w)
    errs = {k: _rel(v.grad, ref_grads[k]) for k, v in model.named_parameters()}
    worst = max(errs, key=errs.get)
    print(f"bf16 mode B={B}: logits {_rel(logits.detach(), ref_logits):.2e}, adjacency {_rel(adj.detach(), ref_adj):.2e}, worst gradient {errs[worst]:.2e} ({worst})")
    assert _rel(adj.detach(), ref_adj) < TOL and _rel(logits.detach(), ref_logits) < 2e-2
    for k, e in errs.items():
        assert e < (2e-1 if (".mean_" in k or ".precision_" in k) else 1e-1), (k, e)
