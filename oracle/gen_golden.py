"""Generate tests/golden/*.npz by running the REFERENCE's own Python (core/corr.py, core/update.py,
core/raft.py, utils/* imported from /root/reference through oracle/refimport.py) on the seeded
synthetic inputs of cer_mvs_b200/synth.py.  TEST INFRASTRUCTURE ONLY; runs only in the build
container.  Usage:  python oracle/gen_golden.py

The CUDA-only ``alt_cuda_corr.forward`` is served by ``cer_oracle.corr_forward`` here; that one
function is pinned separately on the GPU against the compiled reference kernel (oracle/build_ref.py).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import refimport  # noqa: E402
from cer_mvs_b200 import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
H, W, V = 80, 112, 3          # h1 x w1 = 20 x 28; V*px = 1680 (multiple of 16, bilinear_sampler.py:19)


def t(x):
    return torch.from_numpy(np.ascontiguousarray(x))


def load_update_block(ref, sd_np, cascade):
    ub = ref.update.UpdateBlock(cascade=cascade, dim_net=64, dim_inp=64)
    ub.load_state_dict({k: t(v) for k, v in sd_np.items()}, strict=True)
    return ub.eval()


def gen_ops(ref):
    sc = synth.make_scene(H, W, V, seed=1)
    h1, w1 = H // 4, W // 4
    fmaps, poses, K = t(sc["fmaps"]), t(sc["poses"]), t(sc["intrinsics"]).clone()
    K[:, :, :2] /= 4
    ii = torch.zeros(V, dtype=torch.long)
    jj = torch.arange(1, V + 1)
    rs = np.random.RandomState(7)
    out = {}
    true_disp = sc["true_disp"]
    for stage, (D, incre, shift) in enumerate([(64, 0.0025 / 64, True), (44, 0.0025 / 320, False)]):
        if stage == 0:
            disp_in = torch.zeros(1, 1, h1, w1)
        else:
            disp_in = t((true_disp + rs.uniform(-2e-5, 2e-5, true_disp.shape)).astype(np.float32))[None, None]
        with torch.no_grad():
            cb = ref.corr.CorrBlock(fmaps, poses, K, ii, jj, nIncre=D, incre=incre, disps_input=disp_in,
                                    shift=shift, num_levels=3, radius=5, test_mode=True, do_report=False)
            out[f"s{stage}_disp_in"] = disp_in.numpy()
            out[f"s{stage}_origin"] = cb.disps_origin.numpy()
            for l in range(3):
                out[f"s{stage}_pyr{l}"] = cb.corr_pyramid[l].reshape(V * h1 * w1, -1).numpy()
            # lookups: at the truth, at zero (clamped by max(.,0)), far beyond the last hypothesis, random
            zs = {
                "true": true_disp,
                "zero": np.zeros_like(true_disp),
                "far": true_disp + 100 * incre,
                "rand": (cb.disps_origin.numpy()[0, 0, 0] + rs.uniform(-40, 40, true_disp.shape) * incre),
            }
            for name, z in zs.items():
                z = t(z.astype(np.float32))[None, None]
                out[f"s{stage}_z_{name}"] = z.numpy()
                out[f"s{stage}_lookup_{name}"] = cb(z[:, ii]).numpy()
    # projective_transform on its own (one view, 5 hypotheses)
    disps = t(np.linspace(0.0005, 0.0025, 5, dtype=np.float32)).view(1, 1, 5, 1, 1) + torch.zeros(1, 1, 5, h1, w1)
    x1 = ref.pops.projective_transform(poses, disps, K, ii[:1], jj[1:2])
    out["proj_disps"] = disps.numpy()
    out["proj_x1"] = x1.numpy()
    np.savez_compressed(os.path.join(OUT, "ops_corrblock.npz"), **out)
    print("ops_corrblock", {k: v.shape for k, v in out.items() if "pyr0" in k})


def gen_update(ref):
    h1, w1 = H // 4, W // 4
    cascade = [(64, 64, 8), (-1, 320, 8)]
    sd = synth.make_update_weights(seed=2, delta_scale=1.0)
    ub = load_update_block(ref, sd, cascade)
    rs = np.random.RandomState(11)
    net, inp = synth.make_context(h1, w1, seed=2)
    disp = (0.0015 + 2e-4 * synth._smooth_field(rs, 1, h1, w1)).astype(np.float32)[None]
    corr_frames = rs.standard_normal((1, V, 33, h1, w1)).astype(np.float32)
    out = dict(net=net, inp=inp, disp=disp, corr_frames=corr_frames)
    with torch.no_grad():
        out["disp_enc"] = ub.disp_encoder(t(disp)).numpy()
        for stage in (0, 1):
            n2, d2 = ub(t(net), t(inp), t(disp), t(corr_frames), stage)
            out[f"net_out{stage}"] = n2.numpy()
            out[f"delta{stage}"] = d2.numpy()
    np.savez_compressed(os.path.join(OUT, "ops_update.npz"), **out)
    print("ops_update ok")


def run_raft(ref, sc, sd, cascade, scale, autocast_cpu_fp16=False):
    """The reference's RAFT.forward with the two encoders replaced by stubs that hand back the
    synthetic feature / context maps (encoders are outside the hot path)."""
    h1, w1 = H // 4, W // 4
    model = ref.raft.RAFT(cascade=cascade, test_mode=True)
    model.update_block.load_state_dict({k: t(v) for k, v in sd.items()}, strict=True)
    model.eval()
    fm = t(sc["fmaps"])
    pre = t(synth.make_context_pre(h1, w1, seed=sc["seed"]))
    calls = {"i": 0}

    def fnet(x):
        i = calls["i"]
        calls["i"] += 1
        f = fm[:, [i]]
        return f.half() if autocast_cpu_fp16 else f

    def cnet(x):
        return pre.half() if autocast_cpu_fp16 else pre

    class Stub(torch.nn.Module):
        def __init__(self, fn):
            super().__init__()
            self.fn = fn

        def forward(self, x):
            return self.fn(x)

    model.fnet, model.cnet = Stub(fnet), Stub(cnet)
    deltas = []
    model.update_block.register_forward_hook(lambda m, a, o: deltas.append(o[1].float().numpy().copy()))
    images = torch.zeros(1, V + 1, 3, H, W)
    if autocast_cpu_fp16:
        ref.raft.autocast = lambda enabled=True: torch.autocast("cpu", dtype=torch.float16, enabled=enabled)
    with torch.no_grad():
        disp = model(images, t(sc["poses"]).clone(), t(sc["intrinsics"]).clone(),
                     scale=torch.tensor([scale], dtype=torch.float64))
    net = torch.tanh(pre[:, :, :64]).numpy()
    inp = torch.relu(pre[:, :, 64:]).numpy()
    return dict(disp=disp.numpy(), deltas=np.stack(deltas, 0), net=net, inp=inp)


def gen_e2e(ref):
    cascade = [(64, 64, 3), (-1, 320, 3)]
    for name, seed, dscale, dbias, scale in [("trained_like", 3, 0.02, 0.0, 1.0), ("unscaled_oob", 4, 1.0, 0.0, 1.0),
                                             ("drift", 5, 0.1, 0.02, 1.0), ("scaled_pose", 6, 0.1, 0.02, 0.5)]:
        sc = synth.make_scene(H, W, V, seed=seed)
        sc["seed"] = seed
        sd = synth.make_update_weights(seed=seed, delta_scale=dscale, delta_bias=dbias)
        r = run_raft(ref, sc, sd, cascade, scale)
        np.savez_compressed(os.path.join(OUT, f"e2e_fp32_{name}.npz"), seed=seed, delta_scale=dscale,
                            delta_bias=dbias, scale=scale, cascade=np.array(cascade), **r)
        print("e2e", name, "disp mean", float(r["disp"].mean()), "max|delta|", float(np.abs(r["deltas"]).max()))
    # GPU-autocast proxy: the same forward under torch.autocast("cpu", float16)
    try:
        sc = synth.make_scene(H, W, V, seed=5)
        sc["seed"] = 5
        sd = synth.make_update_weights(seed=5, delta_scale=0.1, delta_bias=0.02)
        r = run_raft(ref, sc, sd, cascade, 1.0, autocast_cpu_fp16=True)
        np.savez_compressed(os.path.join(OUT, "e2e_autocast_drift.npz"), seed=5, delta_scale=0.1,
                            delta_bias=0.02, scale=1.0, cascade=np.array(cascade), **r)
        print("e2e autocast ok, disp mean", float(r["disp"].mean()))
    except Exception as e:  # noqa: BLE001
        print("autocast golden not generated:", repr(e))


if __name__ == "__main__":
    torch.manual_seed(0)
    torch.set_num_threads(4)
    os.makedirs(OUT, exist_ok=True)
    ref = refimport.import_reference()
    gen_ops(ref)
    gen_update(ref)
    gen_e2e(ref)
    print({f: os.path.getsize(os.path.join(OUT, f)) for f in sorted(os.listdir(OUT))})
