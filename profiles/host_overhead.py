"""Where does the e2e arm's time go on the host?  Enqueue time (no synchronisation) vs completed time per frame for the
pipeline object with and without the custom-op calls.  Not a bench.py number.

    python profiles/host_overhead.py
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "video-stream-consistency_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)

import torch  # noqa: E402

import bench  # noqa: E402
import synth  # noqa: E402
import vsc_b200 as V  # noqa: E402

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
wl = bench.WORKLOADS["1080p-light"]
W, H, fw, fh = wl["W"], wl["H"], wl["flowW"], wl["flowH"]
ho, hp = bench.make_host_frames(W, H, pin=True)
flf, flb = synth.flows(fw, fh, 3)
d_flf, d_flb = torch.from_numpy(flf).to(dev), torch.from_numpy(flb).to(dev)
sets = bench.make_op_tensors(wl, dev)
st = V.Stabilizer(W, H, 3)
outs = [V.pinned_empty((H, W, 4)) for _ in range(2)]
ext = torch.cuda.ExternalStream(st.compute_stream, device=dev)
N = bench.NFRAMES


def loop(K, ops, push, out):
    st.sync()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(K):
        if ops:
            with torch.cuda.stream(ext):
                bench.run_ops(V, sets)
        st.step(d_flf, d_flb, outs[t & 1] if out else None)
        if push:
            st.push_frame(ho[(t + 2) % N], hp[(t + 2) % N])
        else:
            st.push_frame(ho[(t + 2) % N], hp[(t + 2) % N])
    t1 = time.perf_counter()
    st.sync()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    return (t1 - t0) / K * 1e3, (t2 - t0) / K * 1e3


for t in range(3):
    st.push_frame(ho[t], hp[t])
loop(5, True, True, True)
for ops, out in ((True, True), (False, True), (False, False)):
    enq, tot = loop(40, ops, True, out)
    print(f"ops={ops} out={out}: host enqueue {enq:.3f} ms/frame, completed {tot:.3f} ms/frame", flush=True)
st.close()
