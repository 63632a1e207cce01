"""Fixture for BASELINE configs[0]: the reference's bundled 720p clip (videos/input.mp4) through the stabilizer.

The reference run is `FlowVideoConsistency -c pwcnet-light videos/input.mp4 videos/processed.mp4 out.mp4`
(README.md).  Of its assets only videos/input.mp4 is present (.MISSING_LARGE_BLOBS lists
processed.mp4 and both ONNX models), so the fixture is built as SURVEY 8(c) prescribes:

  stage "decode"  (this container; needs /root/reference and cv2)
      * frames 0..T-1 of videos/input.mp4 decoded with cv2, stored as a JPEG stack (quality 90; the stored frames
        ARE the fixture's original stream -- product and oracle decode the same bytes);
      * per-frame optical flow from a deterministic CPU method (cv2 DIS, medium preset) in the directions the
        reference requests in compute mode -- flowFwd = flow(frame t -> t+1), flowBwd = flow(frame t+1 -> t)
        (videostabilizer.cpp:271-272 runs the model on (1,2) and (2,1)) -- stored at 1/8 resolution as
        int16 in units of 1/8 pixel, values in full-resolution pixels: the test up-samples them with get_bilinear exactly like
        FLOWDOWNSCALE does (flowmodel.cpp:156-165, values not rescaled).  Parity is defined on identical inputs
        AND flow, so the flow source is free;
      * the processed stream is synthesised at load time (tests/config0.py): posterised copy with a seeded
        per-frame gain/offset flicker and noise.
      -> tests/golden/config0_720p.npz
  stage "refgpu"  (GPU box; needs oracle/_ref/libvsc_ref_gpu.so, built here from the unmodified reference)
      * the reference's own CUDA kernels in the doOneStep sequence (oracle/refdrv/ref_gpu.cu) over the T-2
        steps; every 8th pixel of every 8-bit output frame is kept
      -> tests/golden/config0_refgpu.npz   (pins oracle and product at 720p against the reference itself)

    python tests/golden/make_config0_fixture.py decode
    gpurun -- 'python tests/golden/make_config0_fixture.py refgpu gpurun_out/config0_refgpu.npz'
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))

T = 14          # frames 0..13 -> 12 doOneStep calls (current frame 1..12)
FLOW_DOWN = 8
FLOW_Q = 8       # stored flow = round(flow * FLOW_Q) as int16
LATTICE = 8
VIDEO = "/root/reference/videos/input.mp4"


def decode(path):
    import cv2

    cap = cv2.VideoCapture(VIDEO)
    assert cap.isOpened(), VIDEO
    bgr = []
    for _ in range(T):
        ok, f = cap.read()
        assert ok
        bgr.append(f)
    jpg = [cv2.imencode(".jpg", f, [cv2.IMWRITE_JPEG_QUALITY, 90])[1] for f in bgr]
    # the stored bytes are the stream: flows are computed on what the tests will decode
    dec = [cv2.imdecode(j, cv2.IMREAD_COLOR) for j in jpg]
    gray = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in dec]
    cv2.setNumThreads(1)
    dis = cv2.DISOpticalFlow_create(cv2.DISOPTICAL_FLOW_PRESET_MEDIUM)
    H, W = gray[0].shape
    fw, fh = W // FLOW_DOWN, H // FLOW_DOWN
    fwd = np.zeros((T - 2, fh, fw, 2), np.int16)
    bwd = np.zeros((T - 2, fh, fw, 2), np.int16)
    for i, t in enumerate(range(1, T - 1)):
        f = dis.calc(gray[t], gray[t + 1], None)
        b = dis.calc(gray[t + 1], gray[t], None)
        fwd[i] = np.round(cv2.resize(f, (fw, fh), interpolation=cv2.INTER_AREA) * FLOW_Q).astype(np.int16)
        bwd[i] = np.round(cv2.resize(b, (fw, fh), interpolation=cv2.INTER_AREA) * FLOW_Q).astype(np.int16)
    out = {"jpeg_%02d" % i: j.reshape(-1) for i, j in enumerate(jpg)}
    out.update(T=np.int32(T), W=np.int32(W), H=np.int32(H), flow_down=np.int32(FLOW_DOWN), flow_q=np.int32(FLOW_Q), flow_fwd=fwd, flow_bwd=bwd,
               source=np.array("videos/input.mp4 frames 0..%d, cv2 %s" % (T - 1, cv2.__version__)))
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes; |flow| mean", float(np.abs(fwd.astype(np.float32)).mean() / FLOW_Q))


def refgpu(path):
    import torch

    import config0
    from oracle import oracle as O

    assert torch.cuda.is_available() and O.ref_gpu_available()
    c = config0.load()
    W, H = c["W"], c["H"]
    dev = "cuda"
    of = [torch.from_numpy(O.ref_gpu_to_float(x)).to(dev) for x in c["orig8"]]
    pf = [torch.from_numpy(O.ref_gpu_to_float(x)).to(dev) for x in c["proc8"]]
    st = O.RefGpuStepper(W, H, 3, 2)
    last = pf[2].clone()       # preloadProcessedFrames: lastStabilizedFrame <- processedFrames.back()
    lat = np.zeros((len(c["flows"]), (H + LATTICE - 1) // LATTICE, (W + LATTICE - 1) // LATTICE, 4), np.uint8)
    for i, (ff, fb) in enumerate(c["flows"]):
        t = i + 1
        _, rgba = st.step(of[t - 1], of[t], of[t + 1], pf[t - 1], pf[t], pf[t + 1], last,
                           torch.from_numpy(ff).to(dev), torch.from_numpy(fb).to(dev))
        lat[i] = rgba[::LATTICE, ::LATTICE]
    st.close()
    np.savez_compressed(path, lattice=np.int32(LATTICE), rgba=lat,
                        gpu_name=np.array(torch.cuda.get_device_name(0)))
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    stage = sys.argv[1]
    if stage == "decode":
        decode(sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "config0_720p.npz"))
    elif stage == "refgpu":
        refgpu(sys.argv[2] if len(sys.argv) > 2 else os.path.join(HERE, "config0_refgpu.npz"))
    else:
        raise SystemExit(__doc__)
