"""Where a kernel's warps wait, from an `ncu --set full --import-source on` capture: per launch, the stall-reason
totals of the warp-state samples and the instructions that collect the most samples (with their dominant reason).

    python tools/ncu_stalls.py gpurun_out/r01u_dcn_full.ncu-rep [--top 40] [--launch -1]   (some ncu versions list every launch twice on this page: use --launch 0, 2, 4, ...)

Reads the SASS source page (`ncu -i REP --page source --csv --print-source sass`).  This is the reading that showed
the 17-warp DCN kernel to be latency bound (41 % long scoreboard, first FMA after the corner loads) rather than
memory bound -- profiles/r01s_dcn_ab.md.
"""
import argparse
import csv
import io
import subprocess


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('rep')
    ap.add_argument('--top', type=int, default=40)
    ap.add_argument('--launch', type=int, default=-1, help='index of the profiled launch (default: the last one)')
    args = ap.parse_args()
    raw = subprocess.run(['ncu', '-i', args.rep, '--page', 'source', '--csv', '--print-source', 'sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
    if not starts:
        raise SystemExit('no kernels in %s (was it captured with --import-source on?)' % args.rep)
    k = starts[args.launch]
    end = starts[starts.index(k) + 1] if starts.index(k) + 1 < len(starts) else len(rows)
    hdr = rows[k + 1]
    ci = {h: i for i, h in enumerate(hdr)}
    body = [r for r in rows[k + 2:end] if len(r) == len(hdr)]
    stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
    total = sum(int(r[ci['# Samples']]) for r in body)
    print('kernel :', rows[k][1][:120])
    print('samples:', total)
    agg = sorted(((sum(int(r[ci[s]]) for r in body), s) for s in stalls), reverse=True)
    print('reasons:', ', '.join('%s %.1f%%' % (s[6:], 100.0 * v / max(1, total)) for v, s in agg if v))
    print('%8s %6s  %-16s %12s  %s' % ('samples', 'share', 'main reason', 'executed', 'instruction'))
    for r in sorted(body, key=lambda r: -int(r[ci['# Samples']]))[:args.top]:
        n = int(r[ci['# Samples']])
        main_reason = max(stalls, key=lambda s: int(r[ci[s]]))
        print('%8d %5.1f%%  %-16s %12s  %s' % (n, 100.0 * n / max(1, total), main_reason[6:], r[ci['Instructions Executed']],
                                              r[ci['Source']].strip()[:90]))


if __name__ == '__main__':
    main()
