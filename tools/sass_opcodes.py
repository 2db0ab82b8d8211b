"""Per-kernel histogram of the SASS mnemonics that identify the Blackwell / tensor-core paths (cuobjdump -sass on the
built library; runs on the CPU box).  usage: python tools/sass_opcodes.py > profiles/<name>.txt"""
import collections, os, re, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "yolo-litepi_b200", "liblitepi_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UBLKCP", "UTMALDG", "UTMASTG", "HMMA", "LDSM", "LDGSTS", "SYNCS", "UTCBAR", "REDUX", "MUFU", "FFMA", "IMAD", "LDG", "STG", "LDS", "STS", "BAR"]
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        hist.setdefault(kern, collections.Counter())
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
    if m and kern:
        op = m.group(1)
        total[kern] += 1
        for k in KEYS:
            if op.startswith(k):
                hist[kern][k] += 1
                break
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} (sm_100a): instruction counts per kernel; only kernels with tensor-core / TMA / mbarrier opcodes or > 2000 instructions are listed in full")
print(f"{'kernel':70s} {'instr':>7s}  " + " ".join(f"{k:>7s}" for k in KEYS))
for k, h in hist.items():
    print(f"{k[:70]:70s} {total[k]:7d}  " + " ".join(f"{h.get(x, 0):7d}" for x in KEYS))
