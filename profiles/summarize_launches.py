"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import collections
import csv
import sys


def main(path, top=30, last_step=False):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    hdr = rows[hi]
    ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg, cnt = collections.Counter(), collections.Counter()
    body = rows[hi + 2:]
    if last_step:
        # one training step = from the last batched weight-pack launch (start of the forward) to the end of the list
        marks = [i for i, r in enumerate(body) if len(r) > ki and 'weight_pack_batch_kernel' in r[ki]]
        if marks:
            body = body[marks[-1]:]
    for r in body:
        if len(r) <= vi:
            continue
        name = r[ki].split('(')[0].replace('void ', '').replace('<unnamed>::', '')[:80]
        try:
            v = float(r[vi].replace(',', ''))
        except ValueError:
            continue
        agg[name] += v
        cnt[name] += 1
    tot = sum(agg.values())
    print(f"total {tot / 1e6:.2f} ms over {sum(cnt.values())} launches ({path})")
    for n, v in agg.most_common(top):
        print('%6.2f%% %9.3f ms %6d  %s' % (100 * v / tot, v / 1e6, cnt[n], n))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30, last_step=len(sys.argv) > 3 and sys.argv[3] == "last_step")
