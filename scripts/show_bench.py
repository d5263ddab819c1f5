import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "N=%s value=%s busbw/gpu=%s algbw=%s ms=%s roofline=%s ok=%s" % (d.get("n_gpus"), d.get("value"), d.get("busbw_per_gpu_GBps"),
          d.get("algbw_GBps"), d.get("ms_per_step"), (d.get("roofline") or {}).get("frac"), d.get("config", {}).get("correct")))
    print("  e2e=%s" % (d.get("e2e"),))
    print("  clocks=%s launches=%s" % (d.get("clocks"), d.get("gpu_launches")))
    ncl = d.get("nccl") or {}
    if isinstance(ncl, dict) and "error" in ncl:
        print("  nccl error:", ncl["error"])
    nc = {s["bytes"]: s for s in (ncl.get("sweep") if isinstance(ncl, dict) else ncl) or []}
    for s in d.get("sweep") or []:
        n = nc.get(s["bytes"])
        print("  %11d B  %10.2f us  busbw %8.2f GB/s %s  %s" % (s["bytes"], s["us"], s["busbw_GBps"], "" if s.get("correct", True) else "WRONG",
              ("nccl %.2f us %.2f GB/s  x%.2f" % (n["us"], n["busbw_GBps"], n["us"] / s["us"])) if n else ""))
