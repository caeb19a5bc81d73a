#!/usr/bin/env python
"""what does bulk host->device traffic do to (a) back-to-back tiny launches, (b) launch + wait round trips, (c) a streaming
kernel?  Answers which part of a chain of short kernels suffers when the link is busy."""
import threading
import time

import torch


def main():
    dev = "cuda"
    tiny = torch.zeros(32, device=dev)
    big = torch.zeros(256 << 20, dtype=torch.uint8, device=dev)
    big2 = torch.empty_like(big)
    src = torch.empty(128 << 20, dtype=torch.uint8).pin_memory()
    dst = torch.empty(128 << 20, dtype=torch.uint8, device=dev)
    side = torch.cuda.Stream()

    def measure():
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5000):
            tiny.add_(1)
        torch.cuda.synchronize()
        a = (time.perf_counter() - t0) / 5000 * 1e6
        t0 = time.perf_counter()
        for _ in range(1000):
            tiny.add_(1)
            torch.cuda.current_stream().synchronize()
        b = (time.perf_counter() - t0) / 1000 * 1e6
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            big2.copy_(big)
        e1.record()
        torch.cuda.synchronize()
        c = 20 * 2 * big.numel() / (e0.elapsed_time(e1) / 1e3) / 1e9
        return a, b, c

    for traffic in (False, True, False, True):
        stop = [False]

        def pump():
            with torch.cuda.stream(side):
                while not stop[0]:
                    dst.copy_(src, non_blocking=True)
                    side.synchronize()
        th = threading.Thread(target=pump) if traffic else None
        if th:
            th.start()
            time.sleep(0.05)
        a, b, c = measure()
        stop[0] = True
        if th:
            th.join()
        print("background H2D %-5s: %.2f us per queued tiny launch, %.2f us per launch + wait, device copy %.0f GB/s" % (traffic, a, b, c), flush=True)


if __name__ == "__main__":
    main()
