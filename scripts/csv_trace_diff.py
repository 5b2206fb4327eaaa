"""Column-by-column comparison of two SCS per-iteration CSV traces (ScsSettings.log_csv_filename): ours
(scs_b200, host/scs_driver.c log_data_to_csv) against the reference's (src/rw.c:707-861).

    python scripts/csv_trace_diff.py ours.csv reference.csv [--rows N]

Prints, per column, the largest relative difference over the first N rows (default: all common rows) and the row
where it occurs. The reference's LAPACK build writes five spectral-cone names into the header without writing their
values (header under USE_LAPACK, values under USE_SPECTRAL_CONES): columns are therefore matched by POSITION over the
62 columns both files carry; `time` is excluded."""
import sys


def read(path):
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    hdr = [h for h in lines[0].split(",") if h]
    rows = [[v for v in ln.split(",") if v != ""] for ln in lines[1:]]
    return hdr, rows


def compare(a_path, b_path, nrows=None):
    ha, ra = read(a_path)
    hb, rb = read(b_path)
    ncol = min(62, min(len(r) for r in ra + rb))
    n = min(len(ra), len(rb)) if nrows is None else min(nrows, len(ra), len(rb))
    out = {}
    for c in range(ncol):
        name = ha[c] if c < len(ha) else f"col{c}"
        if name == "time":
            continue
        worst, at = 0.0, -1
        for i in range(n):
            x, y = float(ra[i][c]), float(rb[i][c])
            if x != x and y != y:      # NaN in both
                continue
            d = abs(x - y) / max(1.0, abs(y)) if (x == x and y == y) else float("inf")
            if d > worst:
                worst, at = d, i
        out[name] = (worst, at)
    return out, n


if __name__ == "__main__":
    nrows = int(sys.argv[sys.argv.index("--rows") + 1]) if "--rows" in sys.argv else None
    res, n = compare(sys.argv[1], sys.argv[2], nrows)
    print(f"{n} rows compared")
    for k, (w, at) in res.items():
        print(f"{k:36s} {w:10.3e}  (row {at})")
