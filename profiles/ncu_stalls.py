"""Per-instruction stall attribution from an ncu report's source page (needs --import-source on / -lineinfo).

    python profiles/ncu_stalls.py report.ncu-rep kernel-regex [top] [launch-index]

Prints, for the chosen matching launch (default the first): total samples per stall reason, and the `top` SASS instructions with the most
stall samples (with their dominant reason) -- what the warps were waiting on."""
import csv
import subprocess
import sys


def main():
    rep, kre = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre],
                         capture_output=True, text=True).stdout.splitlines()
    # the dump holds one table per launch; take the first
    rows = list(csv.reader(raw))
    start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
    hdr = rows[start[which]]
    end = start[which + 1] - 1 if len(start) > which + 1 else len(rows)
    body = [r for r in rows[start[which] + 1:end] if len(r) == len(hdr)]
    print("#", rows[start[which] - 1][1][:110])
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = {hdr[i]: 0 for i in stall_cols}
    for r in body:
        for i in stall_cols:
            tot[hdr[i]] += int(r[i] or 0)
    all_s = sum(tot.values())
    print("total samples %d: " % all_s + ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / all_s) for k, v in
                                                  sorted(tot.items(), key=lambda kv: -kv[1]) if v))
    si, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
    body.sort(key=lambda r: -int(r[si] or 0))
    for r in body[:top]:
        reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        print("%6s samples  %-64s exec %-9s %s" % (r[si], r[1].strip()[:64], r[ie],
                                                  ", ".join("%s %d" % (n, c) for c, n in reasons if c)))


if __name__ == "__main__":
    main()
