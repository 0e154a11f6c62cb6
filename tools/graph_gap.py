"""How much of bench.py's step time is the boundary between two graph replays:
the same step captured once per graph (what bench.py times) and four steps
(the four input sets) per graph.

    python tools/graph_gap.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    dev = torch.device('cuda:0')
    b, h, w, lt, scale = bench.WORKLOADS['c2']
    hz = bench.Harness(dev, 1, 0, b, h, w, lt, scale, 4)
    for i in range(4):
        hz.step(i)
    torch.cuda.synchronize()
    hz.capture()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(4):
            hz.step(i)
        g4 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g4, stream=side):
            for i in range(4):
                hz.step(i)
    torch.cuda.synchronize()

    def timed(fn, n):
        for _ in range(5):
            fn(0)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            fn(i)
        e1.record()
        e1.synchronize()
        return e0.elapsed_time(e1)

    one = timed(hz.run, 200) / 200
    four = timed(lambda i: g4.replay(), 50) / 200
    print(f'one step per graph : {one * 1e3:.1f} us per step')
    print(f'four steps per graph: {four * 1e3:.1f} us per step')


if __name__ == '__main__':
    main()
