"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list: keep libwts kernels only, aggregate by
(kernel, grid).  usage: summarize_launches.py in.csv [steps] > out.md ; with --filter writes the trimmed CSV."""
import collections
import csv
import sys


def load(path):
    first = open(path).readline()
    if first.startswith('kernel,grid'):                      # already trimmed by --filter
        return [(r[0], r[1], r[2], float(r[3])) for r in list(csv.reader(open(path)))[1:]]
    rows = [l for l in open(path) if l.startswith('"')]
    r = list(csv.reader(rows))
    h = r[0]
    ki, vi, gi, bi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
    out = []
    for x in r[1:]:
        name = x[ki]
        if 'wts::' not in name and not name.startswith(('gemm_', 'void gemm', 'cross_attention', 'layernorm', 'decoder_attention',
                                                         'kv_append', 'decode_select', 'embed_kernel', 'step_inputs',
                                                         'enc_attention', 'dtw', 'prep_', 'to_sb16', 'softmax', 'window_gather',
                                                         'cross_kv', 'frames', 'power', 'logmel', 'gather_rows')):
            continue
        try:
            v = float(x[vi].replace(',', ''))
        except ValueError:
            continue
        out.append((name.split('(')[0].replace('void ', '').replace('wts::', ''), x[gi], x[bi], v))
    return out


def main():
    path = sys.argv[1]
    steps = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith('--') else 1.0
    rows = load(path)
    if '--filter' in sys.argv:
        w = csv.writer(sys.stdout)
        w.writerow(['kernel', 'grid', 'block', 'gpu__time_duration.sum [ns]'])
        for r in rows:
            w.writerow(r)
        return
    agg = collections.defaultdict(lambda: [0, 0.0])
    for name, grid, block, v in rows:
        if name.startswith('to_sb16'):
            continue                    # weight conversion at model load, not part of the measured region
        agg[(name, grid, block)][0] += 1
        agg[(name, grid, block)][1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"| kernel | grid | block | launches | avg us | total us /{steps:g} | share |")
    print("|---|---|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| {k[0]} | {k[1]} | {k[2]} | {v[0]} | {v[1] / v[0] / 1e3:.2f} | {v[1] / 1e3 / steps:.1f} | {100 * v[1] / tot:.1f}% |")
    print(f"\ntotal {tot / 1e6 / steps:.3f} ms")


if __name__ == "__main__":
    main()
