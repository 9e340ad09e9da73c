"""Host<->device copy latency / bandwidth on the GPU box (pinned vs pageable), CUDA-event timed."""
import time
import torch

torch.cuda.init()
dev = torch.device("cuda")
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def t_copy(nbytes, pinned, direction, do_flush=False, reps=5):
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=pinned)
    host.fill_(1)
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    best = 1e9
    wall = 1e9
    for _ in range(reps):
        if do_flush:
            flush.zero_()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        if direction == "h2d":
            d.copy_(host, non_blocking=True)
        else:
            host.copy_(d, non_blocking=True)
        b.record()
        torch.cuda.synchronize()
        wall = min(wall, time.perf_counter() - t0)
        best = min(best, a.elapsed_time(b))
    return best, wall * 1e3


for nbytes in (4096, 370696, 2723840, 64 << 20):
    for pinned in (True, False):
        for direction in ("h2d", "d2h"):
            ev, wall = t_copy(nbytes, pinned, direction)
            evf, wallf = t_copy(nbytes, pinned, direction, do_flush=True)
            print("%9d B %-8s %s: %.3f ms (%.2f GB/s) wall %.3f ms | after L2 flush %.3f ms" %
                  (nbytes, "pinned" if pinned else "pageable", direction, ev, nbytes / ev / 1e6, wall, evf))
