"""Reads tools/kernel_ab.py output (JSON lines) and prints `export CT_<FAMILY>_IMPL=1` for every kernel
family whose new variant (impl 0) errored or missed its parity bound, so the following bench runs on a
known-good configuration. Usage:  eval "$(python tools/pick_safe_env.py gpurun_out/x_ab.jsonl)" """
import json
import sys

BOUND = {"ln_bwd": 5e-3, "attention": 1.2e-2, "gemm": None}
ENV = {"ln_bwd": ["CT_LN_BWD_IMPL"], "attention": ["CT_ATTN_FWD_IMPL", "CT_ATTN_BWD_IMPL"], "gemm": ["CT_GEMM_EPI_IMPL"]}
bad, seen = set(), set()
for path in sys.argv[1:]:
    try:
        lines = open(path).read().splitlines()
    except OSError:
        continue
    for line in lines:
        try:
            r = json.loads(line)
        except ValueError:
            continue
        if r.get("impl") != 0 or r.get("kernel") not in ENV:
            continue
        k = r["kernel"]
        seen.add(k)
        if "error" in r:
            bad.add(k)
            continue
        errs = r["err"] if isinstance(r["err"], dict) else {"e": r["err"]}
        bound = r.get("tol") or BOUND[k]
        for name, v in errs.items():
            if name == "finite":
                if not v:
                    bad.add(k)
            elif not (v == v) or v > bound * (3 if k == "gemm" else 1):
                bad.add(k)
for k in ENV:
    if k in bad or k not in seen:
        for e in ENV[k]:
            print("export %s=1" % e)
print("echo 'kernel families falling back to first-generation variants: %s'" % (sorted(bad | (set(ENV) - seen)) or "none"))
