"""Per-kernel roofline table of one Bloom-560M training step from an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv … tools/step_prof.py`).

Every launch is attributed to its role by its position in the step (the launch order is fixed: tools/step_prof.py
runs the un-graphed step), its algorithmic work is computed from the model shape, and the achieved rate is set against
the measured peaks of MEASURED_PEAKS.json (bf16 TFLOP/s sustained / burst, HBM copy GB/s). ncu times are per launch
with cold caches and unthrottled clocks: use the SHARES and the fractions, not the absolute step time.
    python tools/roofline_table.py gpurun_out/r01w_launches.csv > profiles/r01w_roofline_per_kernel.csv
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B, S, H, NH, F, L, V = 8, 1024, 1024, 16, 4096, 24, 250880
T = B * S


def launches(path):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i
            break
    ki, mi = h.index("Kernel Name"), h.index("Metric Value")
    out = []
    for r in rows[start + 1:]:
        if len(r) <= mi:
            continue
        try:
            us = float(r[mi].replace(",", "")) / 1e3
        except ValueError:
            continue
        out.append((re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("ct::", ""), us))
    return out


def main():
    L_ = launches(sys.argv[1])
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf_peak = float(peaks.get("bf16_tflops", 1667.1))            # burst: kernels timed alone under ncu
    hbm_peak = float(peaks.get("hbm_gbs", 6532.2))
    gflop = lambda m, n, k: 2.0 * m * n * k / 1e9
    roles = collections.OrderedDict()

    def add(role, us, flop_g=None, mbytes=None):
        r = roles.setdefault(role, dict(n=0, us=0.0, gflop=0.0, mb=0.0))
        r["n"] += 1; r["us"] += us
        if flop_g: r["gflop"] += flop_g
        if mbytes: r["mb"] += mbytes

    ce_seen = False
    fwd_gemm = bwd_gemm0 = bwd_gemm1 = 0
    attn_f = 17.18  # GFLOP per layer forward, causal-counted
    for name, us in L_:
        if name.startswith("gemm_tcgen05"):
            kind = int(re.search(r"<(\d)>", name).group(1))
            if not ce_seen:
                if us > 1000:
                    add("LM head forward (8192x250880x1024)", us, gflop(T, V, H))
                else:
                    role = ["QKV forward (K=1024,N=3072)", "h->h forward + f32 residual", "h->4h forward + GELU + saved h",
                            "4h->h forward + f32 residual"][fwd_gemm % 4]
                    g = [gflop(T, 3 * H, H), gflop(T, H, H), gflop(T, F, H), gflop(T, H, F)][fwd_gemm % 4]
                    add(role, us, g); fwd_gemm += 1
            elif us > 1000:
                add("LM head dgrad / wgrad", us, gflop(T, V, H))
            elif kind == 0:
                role = ["4h->h wgrad", "h->4h wgrad", "h->h wgrad (split-K 4)", "QKV wgrad (split-K 3)"][bwd_gemm0 % 4]
                g = [gflop(T, H, F), gflop(T, F, H), gflop(T, H, H), gflop(T, 3 * H, H)][bwd_gemm0 % 4]
                add(role, us, g); bwd_gemm0 += 1
            elif kind == 4:
                add("4h->h dgrad x GELU'(h)", us, gflop(T, F, H))
            else:
                role = ["h->4h dgrad", "h->h dgrad", "QKV dgrad"][bwd_gemm1 % 3]
                g = [gflop(T, H, F), gflop(T, H, H), gflop(T, H, 3 * H)][bwd_gemm1 % 3]
                add(role, us, g); bwd_gemm1 += 1
        elif name.startswith("attn_fwd"):
            add("attention forward", us, attn_f)
        elif name.startswith("attn_bwd"):
            add("attention backward", us, 2.5 * attn_f)
        elif name.startswith("attn_delta"):
            add("attention delta = rowsum(dO*O)", us, mbytes=T * H * 4 / 1e6)
        elif name.startswith("attn_dq_convert"):
            add("attention dQ f32 -> bf16", us, mbytes=T * H * 6 / 1e6)
        elif name.startswith("ln_fwd"):
            add("LayerNorm forward", us, mbytes=T * H * 6 / 1e6)
        elif name.startswith("ln_bwd_fast") or name.startswith("ln_bwd_cta"):
            add("LayerNorm backward (+ residual add, bf16 copy, bias column sums)", us, mbytes=T * H * 14 / 1e6)
        elif name.startswith("ln_bwd_reduce"):
            add("LayerNorm backward: partial reduction", us)
        elif name.startswith("colsum"):
            add("bias gradient column sums", us)
        elif name.startswith("adamw"):
            add("AdamW (flat arena, bf16 shadow)", us, mbytes=559214592 * 30 / 1e6)
        elif name.startswith("ce_fwd"):
            ce_seen = True
            add("cross entropy (loss + dlogits)", us, mbytes=T * V * 4 / 1e6)
        else:
            add("other (embedding, casts, mask prep, CE helpers)", us)
    total = sum(r["us"] for r in roles.values())
    w = csv.writer(sys.stdout)
    w.writerow(["role", "launches", "total_us", "share_pct", "avg_us", "achieved", "unit", "peak", "frac_of_peak"])
    for role, r in sorted(roles.items(), key=lambda kv: -kv[1]["us"]):
        ach = unit = peak = frac = ""
        if r["gflop"]:
            ach = r["gflop"] / r["us"] * 1e3; unit = "TFLOP/s"; peak = tf_peak; frac = ach / peak
        elif r["mb"]:
            ach = r["mb"] / r["us"] * 1e3; unit = "GB/s"; peak = hbm_peak; frac = ach / peak
        w.writerow([role, r["n"], round(r["us"], 1), round(100 * r["us"] / total, 2), round(r["us"] / r["n"], 2),
                    round(ach, 1) if ach != "" else "", unit, peak, round(frac, 3) if frac != "" else ""])
    w.writerow(["TOTAL", sum(r["n"] for r in roles.values()), round(total, 1), 100.0, "", "", "", "", ""])


if __name__ == "__main__":
    main()
