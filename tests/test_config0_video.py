"""BASELINE configs[0]: the reference's bundled 720p clip through >= 10 steps of the recurrence.

Fixture: tests/golden/config0_720p.npz (14 decoded frames of videos/input.mp4 + DIS flows, made by
tests/golden/make_config0_fixture.py) and tests/golden/config0_refgpu.npz (every 8th pixel of the 12 frames the
REFERENCE'S OWN CUDA kernels produce for it on a B200).  Gate: <= 1/255 max-abs on every 8-bit frame
(north_star), for
  * the oracle against the reference-GPU lattice           (CPU, pins the oracle at a BASELINE size),
  * the product (vsc_stabilizer, through the C ABI) against the oracle, every pixel of every frame,
  * the product against the reference-GPU lattice, and against the reference kernels run live beside it.
The stream loop's start-up and end quirks (videostabilizer.cpp:136-164) are part of the sequence.
"""
import os

import numpy as np
import pytest

import config0
import synth

needs_fixture = pytest.mark.skipif(not os.path.exists(config0.FIXTURE), reason="config0_720p.npz missing")
needs_refgpu_fixture = pytest.mark.skipif(not os.path.exists(config0.REFGPU), reason="config0_refgpu.npz not generated")

STEPS = 12


@pytest.fixture(scope="module")
def clip():
    return config0.load()


@pytest.fixture(scope="module")
def oracle_frames(clip, O):
    """[(consisOut f32, rgba8)] for current frames 1..12, lastStabilizedFrame carried in fp32 (:247)"""
    O.use_all_cores()
    of = [O.rgba8_to_f32x3(x) for x in clip["orig8"]]
    pf = [O.rgba8_to_f32x3(x) for x in clip["proc8"]]
    last = pf[2]                      # preloadProcessedFrames: processedFrames.back() (:152)
    outs = []
    for i, (ff, fb) in enumerate(clip["flows"]):
        t = i + 1
        last, rgba = O.do_one_step(of[t - 1], of[t], of[t + 1], pf[t - 1], pf[t], pf[t + 1], last, ff, fb)
        outs.append((last, rgba))
    return outs


@needs_fixture
def test_fixture_is_the_bundled_720p_clip(clip):
    assert (clip["W"], clip["H"], clip["T"]) == (1280, 720, 14)
    assert len(clip["flows"]) == STEPS >= 10
    assert clip["orig8"].shape == (14, 720, 1280, 4) and clip["proc8"].shape == clip["orig8"].shape
    # real motion and a processed stream that actually flickers
    assert max(np.abs(f[0][..., :2]).max() for f in clip["flows"]) > 8.0
    means = clip["proc8"][..., :3].reshape(14, -1).mean(axis=1) - clip["orig8"][..., :3].reshape(14, -1).mean(axis=1)
    assert means.max() - means.min() > 4.0


@needs_fixture
@needs_refgpu_fixture
def test_oracle_matches_reference_gpu_on_config0(clip, oracle_frames):
    """the CPU restatement vs the reference's own kernels (flowconsistency.cu / gpuimage.cu unmodified, sm_100a)"""
    g = np.load(config0.REFGPU)
    L = int(g["lattice"])
    for i, (co, rgba) in enumerate(oracle_frames):
        d = np.abs(rgba[::L, ::L, :3].astype(np.int32) - g["rgba"][i][..., :3].astype(np.int32))
        assert d.max() <= 1, f"frame {i + 1}: {d.max()} grey levels"
        assert (rgba[::L, ::L, 3] == g["rgba"][i][..., 3]).all()


def _product_frames(V, dev, clip):
    import torch

    W, H = clip["W"], clip["H"]
    st = V.Stabilizer(W, H, 3)
    flows = [(torch.from_numpy(ff).to(dev), torch.from_numpy(fb).to(dev)) for ff, fb in clip["flows"]]
    torch.cuda.synchronize()
    for t in range(3):
        st.push_frame(clip["orig8"][t], clip["proc8"][t])
    outs, f32 = [], []
    for i in range(STEPS):
        t = i + 1
        out = np.zeros((H, W, 4), np.uint8)
        st.step(flows[i][0], flows[i][1], out)
        if t + 2 < clip["T"]:
            st.push_frame(clip["orig8"][t + 2], clip["proc8"][t + 2])
        f32.append(st.last_output().cpu().numpy())
        outs.append(out)
    st.sync()
    st.close()
    return outs, f32


@pytest.mark.gpu
@needs_fixture
def test_config0_recurrence_vs_oracle(V, O, dev, clip, oracle_frames):
    outs, f32 = _product_frames(V, dev, clip)
    worst = 0
    for i in range(STEPS):
        d = np.abs(outs[i].astype(np.int32) - oracle_frames[i][1].astype(np.int32))
        worst = max(worst, int(d.max()))
        assert d.max() <= 1, f"frame {i + 1}: {d.max()} grey levels"
        assert (d > 0).mean() < 0.01, f"frame {i + 1}: {(d > 0).mean():.4f} of the bytes differ"
        assert np.abs(f32[i] - oracle_frames[i][0]).max() <= 2e-4, i
    # the whole output sequence as the reference's loop emits it (start-up and end quirks included)
    import torch

    def passthrough(f):
        return V.gpu_to_image(V.image_to_gpu(torch.from_numpy(f).to(dev))).cpu().numpy()

    got = config0.reference_stream_loop(clip["T"], clip["proc8"], lambda t: outs[t - 1])
    ref = config0.reference_stream_loop(clip["T"], clip["proc8"], lambda t: oracle_frames[t - 1][1])
    assert sorted(got) == list(range(clip["T"]))
    for j in (0, clip["T"] - 1):
        assert np.array_equal(got[j], passthrough(clip["proc8"][j]))
        assert np.array_equal(got[j], ref[j])
    assert not np.array_equal(got[1], passthrough(clip["proc8"][1]))   # frame 1 is re-emitted stabilized


@pytest.mark.gpu
@needs_fixture
@needs_refgpu_fixture
def test_config0_recurrence_vs_reference_gpu_fixture(V, dev, clip):
    g = np.load(config0.REFGPU)
    L = int(g["lattice"])
    outs, _ = _product_frames(V, dev, clip)
    for i in range(STEPS):
        d = synth.u8_distance(outs[i][::L, ::L, :3], g["rgba"][i][..., :3])
        assert d.max() <= 1, f"frame {i + 1}: {d.max()} grey levels"


@pytest.mark.gpu
@needs_fixture
def test_config0_recurrence_vs_reference_gpu_live(V, O, dev, clip):
    """the reference's CUDA kernels (oracle/_ref/libvsc_ref_gpu.so) run beside the product on the same B200"""
    if not O.ref_gpu_available():
        pytest.skip("oracle/_ref/libvsc_ref_gpu.so not built")
    import torch

    W, H = clip["W"], clip["H"]
    of = [torch.from_numpy(O.rgba8_to_f32x3(x)).to(dev) for x in clip["orig8"]]
    pf = [torch.from_numpy(O.rgba8_to_f32x3(x)).to(dev) for x in clip["proc8"]]
    ref = O.RefGpuStepper(W, H, 3, 2)
    last = pf[2].clone()
    outs, _ = _product_frames(V, dev, clip)
    for i, (ff, fb) in enumerate(clip["flows"]):
        t = i + 1
        _, rgba = ref.step(of[t - 1], of[t], of[t + 1], pf[t - 1], pf[t], pf[t + 1], last,
                           torch.from_numpy(ff).to(dev), torch.from_numpy(fb).to(dev))
        d = synth.u8_distance(outs[i], rgba)   # mod 256: the 8-bit conversion wraps without a clamp (synth.py)
        assert d.max() <= 1, f"frame {t}: {d.max()} grey levels"
        assert (outs[i][..., 3] == rgba[..., 3]).all()
    ref.close()
