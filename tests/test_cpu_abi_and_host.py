"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/*.h declares, host
logic (packing, partitioning, stage parameters, module substitution), and error paths that need no GPU."""
import ctypes
import os
import re
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "cer_mvs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cer_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from cer_mvs_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in the header but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    L = _lib.lib()
    assert L.cer_abi_version() == 1
    assert L.cer_update_blob_bytes() > 1_000_000
    assert L.cer_update_workspace_bytes(296, 400) > 100_000_000


def test_no_device_is_reported_not_hidden():
    from cer_mvs_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = _lib.lib().cer_device_check()
    assert rc != 0 and b"no CPU fallback" in _lib.lib().cer_last_error()
    import cer_mvs_b200.alt_cuda_corr as acc
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        acc.forward(torch.zeros(1, 2, 2, 4), torch.zeros(1, 2, 2, 4), torch.zeros(1, 1, 2, 2, 2), 0)


def test_argument_validation_without_gpu():
    from cer_mvs_b200 import _lib
    L = _lib.lib()
    assert L.cer_corr_forward_f32(None, None, None, None, 1, 2, 2, 0, 2, 4, 1, 0, None) != 0     # H2 = 0
    assert L.cer_corr_forward_f32(None, None, None, None, 1, 2, 2, 2, 2, 4, 0, 0, None) == 0     # empty: N = 0
    assert L.cer_lookup(None, 1, None, None, 64, 1e-3, 5, 3, None, 4, 4, None) != 0
    assert b"null pointer" in L.cer_last_error()


def test_pack_update_weights_layout():
    """Spot-check the native packer against the layout documented in csrc/update_blob.h."""
    from cer_mvs_b200 import synth
    from cer_mvs_b200.update import pack_update_weights
    sd = synth.make_update_weights(seed=4, fp16_exact=False)
    blob = pack_update_weights(sd)

    def a256(x):
        return (x + 255) & ~255
    off = 0
    w1 = blob[off:off + 48 * 64 * 2].view(np.float16).reshape(48, 64)
    np.testing.assert_array_equal(w1[:33], sd["corr_encoder.0.weight"][:, :, 0, 0].T.astype(np.float16))
    assert not w1[33:].any()
    off = a256(off + 48 * 64 * 2)
    b1 = blob[off:off + 256].view(np.float32)
    np.testing.assert_array_equal(b1, sd["corr_encoder.0.bias"].astype(np.float16).astype(np.float32))
    off = a256(off + 256)
    w2 = blob[off:off + 9 * 64 * 64 * 2].view(np.float16).reshape(9, 64, 64)        # [tap][k][n]
    want = sd["corr_encoder.2.weight"].reshape(64, 64, 9).transpose(2, 1, 0).astype(np.float16)
    np.testing.assert_array_equal(w2, want)
    off = a256(off + 9 * 64 * 64 * 2)
    off = a256(off + 256)                                                            # b2
    wg = blob[off:off + 4 * 9 * 64 * 192 * 2].view(np.float16).reshape(4, 9, 64, 192)
    wz = sd["gru.convz.weight"].reshape(64, 241, 9)
    wq = sd["gru.convq.weight"].reshape(64, 241, 9)
    np.testing.assert_array_equal(wg[0, :, :, 0:64], wz[:, 0:64].transpose(2, 1, 0).astype(np.float16))
    np.testing.assert_array_equal(wg[2, :, :49, 0:64], wz[:, 128:177].transpose(2, 1, 0).astype(np.float16))
    assert not wg[2, :, 49:, :].any() and not wg[0, :, :, 128:].any()                # zero pads
    np.testing.assert_array_equal(wg[3, :, :, 128:192], wq[:, 177:241].transpose(2, 1, 0).astype(np.float16))
    with pytest.raises(RuntimeError, match="expected shape"):
        pack_update_weights({"gru.convz.weight": np.zeros((64, 128, 3, 3), np.float32)})


def test_stage_params_and_partitioning():
    from cer_mvs_b200.dist import replica_range, view_range
    from cer_mvs_b200.hotpath import stage_params
    assert stage_params([(64, 64, 8), (-1, 320, 8)]) == [(64, 0.0025 / 64, 8), (44, 0.0025 / 320, 8)]   # raft.py:76-81
    for n, g in [(10, 1), (10, 2), (10, 4), (10, 8), (15, 4), (7, 8), (7, 3)]:
        parts = [view_range(n, r, g) for r in range(g)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(g - 1))
        sizes = [e - b for b, e in parts]
        assert max(sizes) - min(sizes) <= 1
    assert [view_range(15, r, 4) for r in range(4)] == [(0, 4), (4, 8), (8, 12), (12, 15)]   # 4,4,4,3
    assert replica_range(5, 1, 2) == (3, 5)
    with pytest.raises(ValueError):
        view_range(4, 4, 4)


def test_unit_sharding_covers_every_unit_once():
    """(view, hypothesis) units of the sharded build, incl. more ranks than views (cfg 5: 7 views / 8 GPUs)."""
    from cer_mvs_b200.dist import unit_range, views_of_units
    for V, D, G in [(7, 64, 8), (7, 44, 8), (10, 64, 4), (15, 44, 4), (1, 64, 8), (3, 64, 2), (2, 4, 16)]:
        seen = np.zeros(V * D, np.int32)
        for r in range(G):
            ub, ue = unit_range(V * D, r, G)
            seen[ub:ue] += 1
            vb, ve = views_of_units(ub, ue, D)
            if ue > ub:
                assert vb * D <= ub and ue <= ve * D and ve - vb <= (ue - ub + D - 1) // D + 1
                # the three kernel calls of cer_plan_build_stage_units: tail of the first view, whole views, head of the last
                a, b = ub % D, (ue - 1) % D + 1
                calls = [(vb, vb + 1, a, b)] if ve - vb == 1 else (
                    ([(vb, vb + 1, a, D)] if a else []) + ([(ve - 1, ve, 0, b)] if b < D else []) +
                    [(vb + (1 if a else 0), ve - (1 if b < D else 0), 0, D)])
                cover = np.zeros(V * D, np.int32)
                for v0, v1, d0, d1 in calls:
                    for v in range(v0, v1):
                        cover[v * D + d0:v * D + d1] += 1
                want = np.zeros(V * D, np.int32)
                want[ub:ue] = 1
                assert np.array_equal(cover, want)
            else:
                assert (vb, ve) == (0, 0)
        assert (seen == 1).all()
    assert [unit_range(7 * 64, r, 8) for r in range(8)][0] == (0, 56)          # 448 units -> 56 per GPU


def _gloo_worker(rank, world, port, ret):
    """world_size-2 CPU stand-in for the sharded build: every rank sums its own units of a [V, px, D] table into a
    partial volume, one all_reduce(sum) -- the host logic of DepthHotPath.forward_sharded without a GPU."""
    import torch.distributed as dist
    from cer_mvs_b200.dist import unit_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    V, px, D = 3, 11, 8
    g = torch.Generator().manual_seed(0)
    per_view = torch.randn(V, px, D, generator=g)
    ub, ue = unit_range(V * D, rank, world)
    part = torch.zeros(px, D)
    for u in range(ub, ue):
        part[:, u % D] += per_view[u // D, :, u % D] / V
    dist.all_reduce(part, op=dist.ReduceOp.SUM)
    if rank == 0:
        ret["sum"] = part.numpy()
        ret["want"] = per_view.mean(0).numpy()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_build_host_logic_gloo_world2():
    import torch.multiprocessing as mp
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_gloo_worker, args=(2, 29611, ret), nprocs=2, join=True)
        np.testing.assert_allclose(ret["sum"], ret["want"], rtol=1e-6, atol=1e-6)


def test_update_block_state_dict_matches_reference_keys():
    from cer_mvs_b200.update import ConvGRU, UpdateBlock
    ub = UpdateBlock(cascade=[(64, 64, 8), (-1, 320, 8)], dim_net=64, dim_inp=64)
    keys = list(ub.state_dict().keys())
    want = [f"{m}.{p}" for m in ("corr_encoder.0", "corr_encoder.2", "delta0.0", "delta0.2", "delta1.0", "delta1.2",
                                 "gru.convz", "gru.convr", "gru.convq") for p in ("weight", "bias")]
    assert keys == want                                  # same names and order as core/update.py:58-78
    assert ub.radius == 5 and ub.num_levels == 3         # read by core/raft.py:78-79,90-91
    assert sum(p.numel() for p in ub.parameters()) == 755_778
    assert tuple(ub.gru.convz.weight.shape) == (64, 241, 3, 3)
    assert isinstance(ub.gru, ConvGRU)


def test_install_substitutes_reference_modules():
    import cer_mvs_b200.install as I
    saved = {k: sys.modules.get(k) for k in ("alt_cuda_corr", "core", "core.corr", "core.update", "core.raft")}
    try:
        for name in ("core", "core.corr", "core.update", "core.raft"):
            sys.modules[name] = types.ModuleType(name)
        I.install()
        import cer_mvs_b200.alt_cuda_corr as acc
        from cer_mvs_b200.corr import CorrBlock
        from cer_mvs_b200.update import UpdateBlock
        assert sys.modules["alt_cuda_corr"] is acc
        assert sys.modules["core.corr"].CorrBlock is CorrBlock
        assert sys.modules["core.raft"].UpdateBlock is UpdateBlock and sys.modules["core.raft"].CorrBlock is CorrBlock
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cer_mvs_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "cer_oracle" not in text and "import oracle" not in text, f


def test_three_fma_division_is_correctly_rounded():
    """lookup_common.cuh divides by (W_l - 1) as q0 = x*rcp, r = fma(-q0, d, x), q = fma(r, rcp, q0) with
    rcp = RN(1/d).  Emulate the fp32 FMAs in float64 (all products are exact there) and compare with IEEE fp32
    division for every divisor the kernels can meet (W_l - 1 for D <= 256) over lookup-coordinate-like inputs."""
    rng = np.random.default_rng(0)
    n = 200_000
    xs = np.concatenate([rng.uniform(-6, 70, n), rng.uniform(-6, 6, n), rng.standard_normal(n) * 1e-3,
                         rng.uniform(0, 1e6, n), np.ldexp(rng.uniform(1, 2, n), rng.integers(-20, 60, n))]).astype(np.float32)
    for d in range(1, 256):
        d32 = np.float32(d)
        rcp = np.float32(1.0) / d32
        q0 = (xs * rcp).astype(np.float32)
        r = xs.astype(np.float64) - q0.astype(np.float64) * float(d32)
        r32 = r.astype(np.float32)
        assert np.array_equal(r32.astype(np.float64), r)          # the residual is exactly representable
        q = (q0.astype(np.float64) + r32.astype(np.float64) * float(rcp)).astype(np.float32)
        assert np.array_equal(q, (xs / d32).astype(np.float32)), d


def test_pfm_writer_matches_reference_bytes(tmp_path):
    """prep.write_pfm / readPFM are host file I/O (no GPU): byte-identical to the file the reference's write_pfm wrote
    for the same depth map (tests/golden/ops_io.npz), with and without the pre-flipped fast path."""
    from cer_mvs_b200 import prep
    g = dict(np.load(os.path.join(ROOT, "tests", "golden", "ops_io.npz")))
    p = tmp_path / "d.pfm"
    prep.write_pfm(p, g["depth"])
    assert open(p, "rb").read() == g["pfm_bytes"].tobytes()
    prep.write_pfm(p, np.ascontiguousarray(np.flipud(g["depth"])), flipped=True)
    assert open(p, "rb").read() == g["pfm_bytes"].tobytes()
    np.testing.assert_array_equal(prep.readPFM(p), g["depth"])
    with pytest.raises(Exception):
        prep.write_pfm(p, g["depth"].astype(np.float64))
