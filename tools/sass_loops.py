"""CPU-side static check of a kernel's loops: SASS instruction count of every backward branch's body (cuobjdump -sass),
with a histogram of the instruction classes inside the largest ones.  Usage: python tools/sass_loops.py <lib.so> <name substring>
(the per-element instruction budget of the issue-bound latent kernels is read off here before spending GPU time)."""
import collections, re, subprocess, sys

lib, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
cur, funcs = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        funcs[cur] = []
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m and cur:
        funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
for name, ins in funcs.items():
    if pat not in name:
        continue
    print("==", name, len(ins), "instructions")
    addr = {a: i for i, (a, _) in enumerate(ins)}
    loops = []
    for i, (a, t) in enumerate(ins):
        m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)?(0x[0-9a-f]+)", t)
        if m:
            tgt = int(m.group(1), 16)
            if tgt <= a and tgt in addr:
                loops.append((addr[tgt], i))
    for lo, hi in sorted(loops, key=lambda x: x[0] - x[1])[:4]:
        body = [t for _, t in ins[lo:hi + 1]]
        cls = collections.Counter()
        for t in body:
            op = re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0]
            cls[op] += 1
        print("  loop %#x..%#x: %d instructions" % (ins[lo][0], ins[hi][0], len(body)))
        print("   ", ", ".join("%s %d" % kv for kv in cls.most_common(24)))
