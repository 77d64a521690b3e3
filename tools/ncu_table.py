"""Markdown table + per-element figures from a tools/ncu_summary.py text file of one cfg4 pass.

    python tools/ncu_table.py profiles/r1j_cfg4_all_kernels.txt [elements_in_pass]

Element counts per bucket are the expected cfg4 fractions (bench.py counts them exactly per step:
roofline.per_kernel[*].units_per_step / 1e9) times the elements of the profiled pass."""
import re
import sys

FRACTION = {"normal": 0.750398942, "saddle_far": 0.156741646, "saddle_mid": 0.024410323, "saddle_left": 0.027821899,
            "saddle": 0.011913790, "alternate": 0.007466488, "alternate_left": 0.021246912}
path = sys.argv[1]
n = float(sys.argv[2]) if len(sys.argv) > 2 else 2.0**25
blocks = open(path).read().split("\nkernel: ")[1:]
seen_far = 0
print("| kernel | elements | µs | lanes | FP64 pipe | issue | flops/element | DRAM B/element | TFLOP/s |")
print("|---|---|---|---|---|---|---|---|---|")
for b in blocks:
    name = b.splitlines()[0]
    get = lambda key: float(re.search(rf"{key}\s+([-0-9.e+]+)", b).group(1)) if re.search(rf"{key}\s+([-0-9.e+]+)", b) else float("nan")
    dur = get("duration_ms")
    if " ms" in re.search(r"duration_ms.*", b).group(0):
        dur *= 1000.0
    short, key = name, None
    if "bucket_kernel" in name:
        short, key = "`bucket_kernel`", None
    elif "dense_kernel<5>" in name:
        short, key = "`dense_kernel<normal>`", "normal"
    elif "SaddleT<0>" in name:
        key = "saddle_far" if seen_far == 0 else "saddle_mid"
        short = "`sample_kernel<SaddleFar, fused>`" + ("" if seen_far == 0 else ", 8 ≤ \\|z\\| < 14")
        seen_far += 1
    else:
        which = "setup" if "setup_kernel" in name else "sample"
        for tag, k, label in (("SaddleT<1>", "saddle_left", "SaddleLeft"), ("SaddleT<2>", "saddle", "SaddleFull"),
                              ("AlternateLeft", "alternate_left", "AlternateLeft, fused"), ("Alternate", "alternate", "Alternate"),
                              ("DevroyeT<1>", None, "Devroye pair"), ("DevroyeT<2>", None, "Devroye MSH")):
            if tag in name:
                key, short = k, f"`{which}_kernel<{label}>`"
                break
    el = FRACTION[key] * n if key else (n if "bucket" in name else float("nan"))
    flops = get("fp64_flops_executed")
    dram = (get("dram_read") + get("dram_write")) * 1e6
    fpe = flops / el if el == el else float("nan")
    print(f"| {short} | {el / 1e6:.2f} M | {dur:.0f} | {get('lanes_per_inst'):.1f} | {get('fp64_pipe_pct'):.1f} % | "
          f"{get('issue_active_pct'):.1f} % | {fpe:.0f} | {dram / el:.1f} | {get('fp64_tflops'):.1f} |")
