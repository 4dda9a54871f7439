#!/usr/bin/env python3
"""Do the chains of configs[2] / configs[4] gain from running as P independent lanes (P chain objects of S / P streams,
one CUDA stream each, free-running)?  Every kernel of those chains is one latency-bound wave; lanes that drift apart put
different stages on the SMs at the same time.  Diagnostic, prints ms per step (= one frame of all S streams)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import bench  # noqa: E402
import hoa_cases  # noqa: E402
import hoasp_cases  # noqa: E402
from libmpegh_b200.lib import Lib  # noqa: E402
from oracle import loader as fixtures  # noqa: E402  (layout fixtures only)

lib = Lib().create(0)
GOLD = os.path.join(ROOT, "tests", "golden")
kw = dict(pcm_bits=24, limiter=1, start_delay=240, max_frames_per_call=1)


def run(name, S, P, steps=300, G=2):
    spec = bench.ChainSpec(name, S)
    n = S // P
    handles, lanes = [], []
    if name == "hoa":
        g, gr = np.load(os.path.join(GOLD, "hoasp_golden.npz")), np.load(os.path.join(GOLD, "hoa_ren_golden.npz"))
        k = hoa_cases.key(spec.ren_cfg)
        hsp = lib.hoa_spatial_create(spec.hoasp_cfg, hoasp_cases.golden_tables(g, spec.hoasp_cfg))
        hren = lib.hoa_renderer_create(gr[k + "_M"], gr[k + "_nfc"])
        handles = [(lib.hoa_renderer_destroy, hren), (lib.hoa_spatial_destroy, hsp)]
        sides = spec.hoa_sides(n, 4)
        side_h = [np.ascontiguousarray(sides[i]) for i in range(4)]
        side_key, sigma = "hoa_side", 2000.0
    else:
        lay = lib.obj_layout_create(fixtures.obj_layout_golden(19))
        handles = [(lib.obj_layout_destroy, lay)]
        p = spec.obj_params(n, 2)
        side_h = [lib.obj_side_from_polar(p[i]) for i in range(2)]
        side_key, sigma = "obj_side", 3000.0
    side_d = [torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).cuda() for a in side_h]
    for lane in range(P):
        if name == "hoa":
            chain = lib.chain_open(n, n_hoa_transport=16, hoa_spatial=hsp, hoa_renderer=hren, **kw)
        else:
            chain = lib.chain_open(n, n_obj=32, obj_layout=lay, **kw)
        core = [torch.randn(1, n, spec.n_in, 1024, device="cuda") * sigma for _ in range(G)]
        pcm = torch.empty(n, chain.record_bytes, dtype=torch.uint8, device="cuda")
        nb = torch.zeros(n, dtype=torch.int32, device="cuda")
        frames = [chain.frame(core=core[i % G].data_ptr(), pcm=pcm.data_ptr(), pcm_bytes=nb.data_ptr(),
                              **{side_key: side_d[i % len(side_d)].data_ptr()}) for i in range(4)]
        lanes.append((chain, torch.cuda.Stream(), frames, core, pcm, nb))
    torch.cuda.synchronize()
    main = torch.cuda.current_stream()

    def go(k0, k):
        for i in range(k0, k0 + k):
            for chain, st, frames, *_ in lanes:
                chain.run_dev(frames[i % len(frames)], 1, st.cuda_stream)

    go(0, 20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for _, st, *_ in lanes:
        st.wait_event(e0)
    go(20, steps)
    for _, st, *_ in lanes:
        ev = torch.cuda.Event()
        ev.record(st)
        main.wait_event(ev)
    e1.record(main)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    print(f"{name:8s} S={S} lanes={P}: {ms * 1000:8.1f} us per frame of all streams", flush=True)
    for chain, *_ in lanes:
        chain.close()
    for fn, h in handles:
        fn(h)


for name, S in (("objects", 512), ("hoa", 1024)):
    for P in (1, 2, 4, 8):
        run(name, S, P)
# more streams in flight per GPU: the same question at the north-star stream count
for name, S in (("objects", 4096), ("hoa", 4096)):
    for P in (1, 4):
        run(name, S, P, steps=60)
lib.close()
