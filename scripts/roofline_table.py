"""Per-shape roofline classification of the udt_igemm launches of one UNet CFG step: for every distinct problem the
algorithmic FLOPs, the bytes that must cross HBM when nothing is cached (fp16 A once + weights once + residual + output),
the time each roofline allows (MEASURED_PEAKS.json: sustained tensor TFLOP/s, HBM GB/s) and the measured time ->
which roofline binds and what fraction of it is reached.  Input: a step profile written by scripts/profile_step_graph.py.
usage: python scripts/roofline_table.py profiles/r01c_step_profile_b4.txt > profiles/r02_igemm_roofline_b4.txt"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path = sys.argv[1]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    tf, tfb, bw = peaks.get("bf16_tflops_sustained", 1382.4), peaks.get("bf16_tflops", 1667.4), peaks.get("hbm_gbs", 6547.5)
    rows = []
    for line in open(path):
        m = re.match(r"\s*([\d.]+) ms\s+x(\d+)\s+([\d.]+) us\s+udt_igemm\s+\((\d+), (\d+), (\d+), '([^']*)', '([^']*)', (\d)\)", line)
        if not m:
            continue
        total_ms, calls, us, M, N, K, tag, dims, act = m.groups()
        calls, us, M, N, K, act = int(calls), float(us), int(M), int(N), int(K), int(act)
        n_log = N // 2 if act == 2 else N
        flop = 2.0 * M * N * K
        # A: a 3x3 / 2x2 window re-reads each pixel from L2, HBM sees the tensor once; segments add their own tensors
        a_bytes = 0
        for seg in tag.split("+"):
            taps, c = seg.split("x")
            c = int(re.match(r"\d+", c).group())
            a_bytes += M * c * 2
        w_bytes = N * K * 2
        o_bytes = M * n_log * 2
        byts = a_bytes + w_bytes + o_bytes          # (+ residual where present: not in the profile key; noted below)
        t_tensor, t_hbm = flop / (tf * 1e12) * 1e6, byts / (bw * 1e9) * 1e6
        bound = "tensor" if t_tensor >= t_hbm else "hbm"
        rows.append((calls * us, calls, us, M, N, K, tag, flop / byts, t_tensor, t_hbm, bound, max(t_tensor, t_hbm) / us))
    rows.sort(reverse=True)
    print(f"udt_igemm launches of one UNet CFG step ({os.path.basename(path)}); peaks: tensor {tf:.0f} TFLOP/s sustained "
          f"({tfb:.0f} burst), HBM {bw:.0f} GB/s; ridge {tf * 1e3 / bw:.0f} FLOP/B")
    print("bytes = fp16 A once + weights once + output once (a residual operand adds M*N*2 more: the reached fraction of "
          "those rows is a lower bound); time = each shape replayed alone in a CUDA graph (burst clocks)")
    print(f"{'ms/step':>8} {'x':>3} {'us':>7}  {'M':>6} {'N':>5} {'K':>6}  {'FLOP/B':>6} {'t_tensor':>8} {'t_hbm':>6}  bound   reached  segments")
    tot = tot_floor = 0.0
    for r in rows:
        ms, calls, us, M, N, K, tag, inten, tt, th, bound, frac = r
        tot += ms
        tot_floor += calls * max(tt, th)
        print(f"{ms / 1e3:8.4f} {calls:3d} {us:7.2f}  {M:6d} {N:5d} {K:6d}  {inten:6.0f} {tt:8.2f} {th:6.2f}  {bound:6s}  {frac:6.2f}   {tag}")
    print(f"sum {tot / 1e3:.3f} ms; sum of per-shape roofline floors {tot_floor / 1e3:.3f} ms -> the igemm launches reach "
          f"{tot_floor / tot:.2f} of their own (tensor- or HBM-) rooflines")


if __name__ == "__main__":
    main()
