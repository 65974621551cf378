"""do H2D and D2H copies of pipeline-chunk size overlap on this box? (torch, pinned buffers, two streams)"""
import time, torch
n_chunk, mb_up, mb_down = 64, 2.1, 3.7
up = [torch.empty(int(mb_up * 2**20), dtype=torch.uint8).pin_memory() for _ in range(4)]
dn = [torch.empty(int(mb_down * 2**20), dtype=torch.uint8).pin_memory() for _ in range(4)]
d_up = [torch.empty_like(u, device="cuda") for u in up]
d_dn = [torch.empty_like(u, device="cuda") for u in dn]
s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
def run(do_up, do_dn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n_chunk):
        if do_up:
            with torch.cuda.stream(s_up): d_up[i % 4].copy_(up[i % 4], non_blocking=True)
        if do_dn:
            with torch.cuda.stream(s_dn): dn[i % 4].copy_(d_dn[i % 4], non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n_chunk * 1e6
for _ in range(2):
    a, b, c = run(True, False), run(False, True), run(True, True)
    print("per chunk: H2D %.1f MB alone %.1f us (%.1f GB/s) | D2H %.1f MB alone %.1f us (%.1f GB/s) | both %.1f us" % (mb_up, a, mb_up * 2**20 / a / 1e3, mb_down, b, mb_down * 2**20 / b / 1e3, c))
