"""Summarise an `ncu --page source --csv` export: sample share per region of the kernel + the hottest instructions."""
import csv
import sys


def main(path, nbins=40, top=25):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ia, isrc, isamp, iexec = hdr.index("Address"), hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    body = rows[2:]
    total = sum(int(r[isamp] or 0) for r in body)
    print("instructions", len(body), "total samples", total)
    n = len(body)
    step = max(1, n // nbins)
    for b in range(0, n, step):
        chunk = body[b:b + step]
        s = sum(int(r[isamp] or 0) for r in chunk)
        ex = sum(int(r[iexec] or 0) for r in chunk)
        ops = {}
        for r in chunk:
            op = r[isrc].split()[0] if not r[isrc].strip().startswith("@") else r[isrc].split()[1]
            ops[op] = ops.get(op, 0) + 1
        topops = sorted(ops.items(), key=lambda kv: -kv[1])[:4]
        print(f"[{b:5d}-{b + len(chunk):5d}] samples {100.0 * s / max(total, 1):5.1f}%  exec {ex:12d}  {topops}")
    print("--- hottest instructions")
    for r in sorted(body, key=lambda r: -int(r[isamp] or 0))[:top]:
        st = sorted(((hdr[i], int(r[i] or 0)) for i in stall_cols), key=lambda kv: -kv[1])[:2]
        print(f"{100.0 * int(r[isamp]) / max(total, 1):5.2f}%  {body.index(r):5d}  {r[isrc].strip()[:70]:70s} {st}")


if __name__ == "__main__":
    main(sys.argv[1])
