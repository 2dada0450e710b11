"""CPU-side checks of the boundary: the shared library loads, exports every symbol include/vqa_b200.h declares,
and the host layer refuses to run without CUDA (no silent fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vqa_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vqa_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    from vqa_b200 import _cabi
    assert sorted(_cabi.EXPORTS) == _declared_symbols()


def test_library_loads_and_exports_every_declared_symbol():
    from vqa_b200 import _cabi
    assert os.path.exists(_cabi.LIB_PATH), "run __graft_entry__.build() first"
    lib = _cabi.load()
    for name in _declared_symbols():
        assert hasattr(lib, name), name
    assert lib.vqa_abi_version() == _cabi.ABI_VERSION
    assert lib.vqa_graphconv_edge_blocks(2, 36, 16) == (2 * 36 * 16 + 255) // 256


def test_argument_errors_are_reported_without_a_gpu():
    from vqa_b200 import _cabi
    with pytest.raises(_cabi.VqaKernelError) as e:     # null operands are rejected before any CUDA call
        _cabi.call("vqa_gemm_f32", None, 4, 0, None, 4, 0, None, 4, 1, 1, 1, None, None, 0, 1, None, 0, 1.0, 0, 0, 1, 0, None)
    assert "null operand" in str(e.value)
    with pytest.raises(_cabi.VqaKernelError):
        _cabi.call("vqa_dropout_f32", None, None, 10, 0.5, 1, 1, None, None)


def test_no_cpu_fallback():
    from vqa_b200 import kernels as kn
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kn.gemm(torch.randn(8, 32), torch.randn(8, 32))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        kn.adjacency_topk_fwd(torch.randn(1, 4, 8), 2)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "vqa-project_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, flags=re.M), os.path.join(dirpath, f)


def test_drop_in_state_dict_contract():
    """Same keys, order and shapes as the reference's state_dict (golden file holds the reference's)."""
    import numpy as np
    from conftest import load_golden
    from vqa_b200.synthetic import WORKLOADS, make_wemb
    import sparse_graph_model as M
    g = load_golden("tiny")
    w = WORKLOADS["tiny"]
    torch.manual_seed(1000)
    model = M.Model(pretrained_wemb=make_wemb(w), **w.model_kwargs())
    sd = model.state_dict()
    ref_keys = [k[6:] for k in g if k.startswith("param.")]
    assert list(sd.keys()) == ref_keys
    for k in ref_keys:
        assert tuple(sd[k].shape) == g["param." + k].shape, k
    # same constructor RNG stream as the reference: identical seeded init (the golden script clamps two tensors afterwards)
    same = [k for k in ref_keys if np.array_equal(sd[k].numpy(), g["param." + k])]
    assert len(same) >= len(ref_keys) - 4
    model.load_state_dict({k: torch.from_numpy(g["param." + k]) for k in ref_keys})   # reference checkpoints load
    lst = model.graph_convolution_1.conv_weight_list()
    assert all(x.data_ptr() == lst[0].data_ptr() + i * x.numel() * 4 for i, x in enumerate(lst))


def test_conv_weights_are_rejoined_by_module_conversions():
    """``.to()/.double()`` split the per-kernel weights into storages of their own; the layer re-joins them inside ``_apply`` so the
    addresses a gradient reducer / optimiser records right after ``model.cuda()`` stay valid."""
    import torch
    import layers
    gc = layers.NeighbourhoodGraphConvolution(12, 8, 4, 2)

    def consecutive(m):
        ws = [lin.weight for lin in m.conv_weights]
        step = ws[0].numel() * ws[0].element_size()
        return all(w.data_ptr() == ws[0].data_ptr() + i * step for i, w in enumerate(ws))

    assert consecutive(gc)
    before = [w.detach().clone() for w in (lin.weight for lin in gc.conv_weights)]
    gc = gc.double()
    assert consecutive(gc) and gc.conv_weights[0].weight.dtype == torch.float64
    gc = torch.nn.Sequential(gc).float()[0]               # through a parent module's recursion
    assert consecutive(gc)
    assert all(torch.equal(a, lin.weight) for a, lin in zip(before, gc.conv_weights))


def test_all_steps_row_gate_marks_the_tiles_of_ended_questions():
    """ops._all_steps_gate: the gate of a product over all T*B time-major rows (the GRU's input projection and its data gradient):
    128-row tile (t, j) is skipped (entry <= 0) exactly when every question of tile j is shorter than t + 1."""
    from vqa_b200 import ops
    T, B = 6, 384
    tile_len = torch.tensor([6, 3, 1], dtype=torch.int32)                 # longest question per 128-row tile (batches are length-sorted)
    gate, t0, tag = ops._all_steps_gate(tile_len, T, B)
    assert t0 == 0 and tag == "all_steps" and gate.dtype == torch.int32 and gate.shape == (T * 3,)
    live = (gate > 0).view(T, 3)
    for t in range(T):
        for j in range(3):
            assert bool(live[t, j]) == (int(tile_len[j]) > t)
    assert ops._all_steps_gate(torch.tensor([5], dtype=torch.int32), T, 100) is None      # a step is not a whole number of tiles: no gating
