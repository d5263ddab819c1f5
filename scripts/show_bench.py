import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f, "N=%s value=%s ms=%s ok=%s e2e=%s clocks=%s" % (d.get("n_gpus"), d.get("value"), d.get("ms_per_step"), d.get("config", {}).get("correct"), d.get("e2e"), d.get("clocks")))
    nc = {s["bytes"]: s for s in (d.get("nccl") or [])}
    for s in d.get("sweep") or []:
        n = nc.get(s["bytes"])
        print("  %11d B  %10.2f us  busbw %8.2f GB/s   %s" % (s["bytes"], s["us"], s["busbw_GBps"], ("nccl %.2f us %.2f GB/s" % (n["us"], n["busbw_GBps"])) if n else ""))
