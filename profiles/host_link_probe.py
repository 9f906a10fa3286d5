#!/usr/bin/env python
"""Aggregate device->host (and host->device) copy bandwidth of this box with 1 .. N GPUs copying at the same time.

    python profiles/host_link_probe.py [--mb 1024] [--reps 5]

One pinned host buffer and one device buffer per GPU, one copy stream per GPU, all copies of a round issued before any
is waited for; the time of a round is the wall clock from the first issue to the last completion.  This is the ceiling
of every end-to-end number whose result is a host array: a product of B result bytes cannot leave N GPUs faster than
B / d2h_gbs[N].  Prints one JSON line.
"""
import argparse
import json
import time

import torch


def measure(devs, mb, reps, direction):
    n = mb << 20
    host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in devs]
    dev = [torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d) for d in devs]
    streams = [torch.cuda.Stream(device=d) for d in devs]

    def round_():
        for h, g, s, d in zip(host, dev, streams, devs):
            with torch.cuda.device(d), torch.cuda.stream(s):
                if direction == "d2h":
                    h.copy_(g, non_blocking=True)
                else:
                    g.copy_(h, non_blocking=True)
        for s in streams:
            s.synchronize()

    round_()
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        round_()
        dt = time.perf_counter() - t0
        best = dt if best is None or dt < best else best
    return len(devs) * n / best / 1e9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    ngpu = torch.cuda.device_count()
    out = {"gpus": ngpu, "buffer_mb_per_gpu": a.mb, "d2h_gbs": {}, "h2d_gbs": {}}
    counts = sorted({c for c in (1, 2, 4, 8, ngpu) if c <= ngpu})
    for c in counts:
        devs = list(range(c))
        out["d2h_gbs"][str(c)] = round(measure(devs, a.mb, a.reps, "d2h"), 2)
        out["h2d_gbs"][str(c)] = round(measure(devs, a.mb, a.reps, "h2d"), 2)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
