"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/mgproto_b200.h declares (no compute calls here), the ctypes table matches the header,
and the product path refuses to run without a CUDA device instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mgproto_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mgp_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported():
    from mgproto_b200 import _lib
    names = _declared()
    assert len(names) >= 20
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "library does not export %s" % n
    assert sorted(_lib.SIGNATURES) == names, set(_lib.SIGNATURES) ^ set(names)


def test_ctypes_signatures_match_header_prototypes():
    """Every prototype in include/mgproto_b200.h, argument by argument (pointer / int / float / double / size_t),
    against the ctypes table: a float bound where the header says double would silently corrupt the call."""
    import ctypes as C
    from mgproto_b200 import _lib
    src = open(os.path.join(ROOT, "include", "mgproto_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|const char\s*\*)\s+(mgp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src)
    assert len(protos) == len(_lib.SIGNATURES)

    def kind(arg):
        arg = arg.strip()
        if "*" in arg:
            return "ptr"
        for k in ("double", "float", "size_t", "int"):
            if re.search(r"\b%s\b" % k, arg):
                return k
        raise AssertionError(arg)

    cmap = {C.c_void_p: "ptr", C.c_char_p: "ptr", C.c_int: "int", C.c_float: "float", C.c_double: "double",
            C.c_size_t: "size_t"}
    for name, args in protos:
        want = [] if args.strip() in ("", "void") else [kind(a) for a in args.split(",")]
        got = [cmap[t] for t in _lib.SIGNATURES[name][1]]
        assert got == want, (name, got, want)


def test_abi_version_and_errors():
    from mgproto_b200 import _lib
    lib = _lib.load()
    assert lib.mgp_abi_version() == 2
    assert b"invalid" in lib.mgp_error_string(-1)
    assert b"supported" in lib.mgp_error_string(-2)
    # argument validation happens before any CUDA call, so it is testable without a GPU
    assert lib.mgp_normalize_fwd(None, None, None, None, 1, 1, 1, None) == -1
    assert lib.mgp_em_stat_stride(10, 128, 0) == 10 + 1280 + 1
    assert lib.mgp_em_stat_stride(10, 128, 1) == 10 + 2560 + 1
    with pytest.raises(_lib.MGProtoLibraryError):
        _lib.check(-2, "x")


def test_no_cpu_fallback():
    import mgproto_b200 as M
    net = M.construct_MGProto("resnet18", pretrained=False, prototype_shape=(20, 16, 1, 1), num_classes=4,
                              add_on_layers_type="regular", mem_capacity=8, mine_K=3)
    x = torch.randn(1, 3, 32, 32)
    with pytest.raises(RuntimeError, match="CUDA"):
        net(x, None)
    with pytest.raises(RuntimeError, match="CUDA"):
        net.compute_log_prob(torch.randn(8, 16))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "mgproto_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            assert "oracle" not in open(os.path.join(pkg, fn)).read().replace("oracle/_ref", ""), fn


def test_reference_surface():
    """Constructor / attributes / state-dict keys the reference's callers rely on."""
    import mgproto_b200 as M
    net = M.construct_MGProto("resnet18", pretrained=False, prototype_shape=(12, 16, 1, 1), num_classes=4,
                              add_on_layers_type="bottleneck", sz_embedding=8, mem_capacity=6, mine_K=3)
    for a in ("queue", "capacity_pc", "num_classes", "num_prototypes", "num_prototypes_per_class",
              "iteration_counter", "update_interval", "prototype_means", "prototype_covs", "prototype_class_identity",
              "prototype_shape", "last_layer", "features", "add_on_layers", "prototype_optimizer", "img_size",
              "mine_T", "num_em_loop", "alpha", "tau", "memory_updated_cls"):
        assert hasattr(net, a), a
    for m in ("forward", "push_forward", "update_GMM", "compute_log_prob", "_e_step", "_estimate_log_prob", "_m_step",
              "_score", "prune_prototypes_topM", "set_last_layer_incorrect_connection", "conv_features"):
        assert callable(getattr(net, m)), m
    assert tuple(net.prototype_means.shape) == (4, 3, 16) and net.prototype_means.requires_grad
    assert not net.prototype_covs.requires_grad and not net.last_layer.weight.requires_grad
    assert tuple(net.last_layer.weight.shape) == (4, 12)
    w = net.last_layer.weight
    assert torch.allclose(w.sum(1), torch.ones(4)) and float(w[0, 3:].abs().sum()) == 0.0      # KA3 init
    assert torch.allclose(net.prototype_means.norm(dim=2), torch.ones(4, 3), atol=1e-6)
    sd = net.state_dict()
    for k in ["prototype_means", "prototype_covs", "iteration_counter", "last_layer.weight", "queue.mem_len"] + \
             ["queue.cls%d" % i for i in range(4)]:
        assert k in sd, k
    assert sd["queue.cls0"].shape == (6, 16) and sd["queue.mem_len"].dtype == torch.int64
    assert not any(k.startswith("queue.") and k.split(".")[1] in ("bank", "head", "updated") for k in sd)
    net.load_state_dict(sd)
    # pruning keeps >= 1 prototype per class and zeroes the rest (ref model.py:467-482)
    net.last_layer.weight.data[0, :3] = torch.tensor([0.5, 0.3, 0.2])
    net.prune_prototypes_topM(top_M=2)
    assert float(net.last_layer.weight[0, 2]) == 0.0 and float(net.last_layer.weight[0, 1]) > 0


def test_memory_bank_cpu_bookkeeping():
    """Ring <-> reference-order conversion and explicit push() (no kernels involved on CPU)."""
    from mgproto_b200.memory import MemoryBank
    import numpy as np
    from oracle import mgproto_oracle as O
    mb = MemoryBank(3, 4, 3 * 5)
    ob = O.MemoryBankOracle(3, 4, 5)
    g = torch.Generator().manual_seed(0)
    for _ in range(9):
        n = int(torch.randint(1, 5, (1,), generator=g))
        feat = torch.randn(n, 4, generator=g)
        lab = torch.randint(0, 3, (n,), generator=g)
        mb.push(feat, lab)
        for c in torch.unique(lab).tolist():
            ob.push(c, feat[lab == c].numpy())
        lin = mb.linear().numpy()
        assert mb.mem_len.tolist() == ob.mem_len.tolist()
        for c in range(3):
            np.testing.assert_array_equal(lin[c, :ob.mem_len[c]], ob.data[c, :ob.mem_len[c]])
    data, labels = mb.pull_all()
    assert data.shape[0] == int(mb.mem_len.sum()) and labels.tolist() == sorted(labels.tolist())


def test_header_constants_match_binding():
    """#define MGP_* values in include/mgproto_b200.h == the Python mirror in _lib.py (layouts, math modes, errors)."""
    import re
    from mgproto_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mgproto_b200.h")).read()
    defs = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(MGP_[A-Z0-9_]+)\s+(-?\d+)\b", hdr)}
    for name in ("MGP_OUT_LOGP_NP", "MGP_OUT_LOGP_BPHW", "MGP_OUT_NEGP_BPHW", "MGP_OUT_TOP1_BP", "MGP_MATH_FP32",
                 "MGP_MATH_TC", "MGP_MATH_AUTO", "MGP_MATH_TC_REUSE", "MGP_MATH_TC_ISO", "MGP_MATH_TC_ISO_REUSE"):
        assert name in defs, name
        assert getattr(_lib, name) == defs[name], name
