"""hot SASS lines of an ncu report: python scripts/ncu_hot.py rep [min_frac]"""
import csv, subprocess, sys
rep = sys.argv[1]
frac = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
ts = sum(int(r[ix['# Samples']]) for r in data); ti = sum(int(r[ix['Instructions Executed']]) for r in data)
print("total samples", ts, "total warp instr", ti)
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
for k, r in enumerate(data):
    sm = int(r[ix['# Samples']])
    src = r[ix['Source']].strip()
    if sm >= ts * frac or any(x in src for x in ("BAR.", "SYNCS", "UBLKCP")):
        top = sorted(((int(r[ix[c]]), c[6:]) for c in stalls), reverse=True)[:2]
        print(k, src[:64].ljust(64), r[ix['Instructions Executed']].rjust(9), str(sm).rjust(6), r[ix['Avg. Threads Executed']].rjust(3), top)
