"""H2D copy of pinned buffers on this box: time vs size for one copy and for 10 back-to-back copies (events on the copy's own
stream), to separate per-copy latency from bandwidth.  The C2 batch is 7.08 MB."""
import json, torch
dev = torch.device("cuda:0")
out = {}
for mb in (0.25, 1, 2, 4, 7.08, 16, 64):
    n = int(mb * (1 << 20)) // 4
    h = torch.rand(n).pin_memory(); d = torch.empty(n, device=dev)
    for reps, tag in ((1, "single"), (10, "x10")):
        ts = []
        for r in range(8):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps): d.copy_(h, non_blocking=True)
            e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) / reps)
        t = min(ts[2:])
        out[f"{mb}MB_{tag}"] = {"ms_per_copy": round(t, 4), "GBs": round(n * 4 / (t * 1e-3) / 1e9, 1)}
print(json.dumps(out))
