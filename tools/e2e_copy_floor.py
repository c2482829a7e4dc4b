"""Dev tool (GPU box): the copy-only floor of bench.py's single-GPU e2e loop - the same byte counts and stream pattern as the
pipelined composers (8 x 36 MB up, then 181 MB + 60 MB down, per step; `depth` steps in flight on `depth` streams), without
any kernel.  Variant `split`: all uploads on one stream and all downloads on another, chained by events.

  python tools/e2e_copy_floor.py [--steps 30] [--depth 3]
"""
import argparse

import torch


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--depth", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n_img, img_b = 8, 4000 * 3000 * 3
    out8_b, outm_b = 20912 * 2881 * 3, 20912 * 2881
    h_imgs = [torch.empty(img_b, dtype=torch.uint8).pin_memory() for _ in range(n_img)]
    slots = []
    for _ in range(a.depth):
        slots.append(dict(d_imgs=[torch.empty(img_b, dtype=torch.uint8, device=dev) for _ in range(n_img)],
                          d8=torch.zeros(out8_b, dtype=torch.uint8, device=dev), dm=torch.zeros(outm_b, dtype=torch.uint8, device=dev),
                          h8=torch.empty(out8_b, dtype=torch.uint8).pin_memory(), hm=torch.empty(outm_b, dtype=torch.uint8).pin_memory()))
    streams = [torch.cuda.Stream(dev) for _ in range(a.depth)]
    up, down = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def same_stream(steps):
        for k in range(steps):
            s, sl = streams[k % a.depth], slots[k % a.depth]
            with torch.cuda.stream(s):
                for i in range(n_img):
                    sl["d_imgs"][i].copy_(h_imgs[i], non_blocking=True)
                sl["h8"].copy_(sl["d8"], non_blocking=True)
                sl["hm"].copy_(sl["dm"], non_blocking=True)

    def split(steps):
        evs = []
        for k in range(steps):
            sl = slots[k % a.depth]
            with torch.cuda.stream(up):
                if k >= a.depth:
                    up.wait_event(evs[k - a.depth][1])  # the slot's previous download
                for i in range(n_img):
                    sl["d_imgs"][i].copy_(h_imgs[i], non_blocking=True)
                e_up = torch.cuda.Event()
                e_up.record(up)
            with torch.cuda.stream(down):
                down.wait_event(e_up)
                sl["h8"].copy_(sl["d8"], non_blocking=True)
                sl["hm"].copy_(sl["dm"], non_blocking=True)
                e_dn = torch.cuda.Event()
                e_dn.record(down)
            evs.append((e_up, e_dn))

    for name, fn in (("same stream per step", same_stream), ("upload stream + download stream", split)):
        fn(2 * a.depth)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams + [up, down]:
            s.wait_event(e0)
        fn(a.steps)
        for s in streams + [up, down]:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        print(f"{name}: {ms:.3f} ms/step  up {n_img * img_b / ms / 1e6:.1f} GB/s  down {(out8_b + outm_b) / ms / 1e6:.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
