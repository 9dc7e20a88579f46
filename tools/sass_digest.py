"""Mnemonic histogram of every kernel in libmcut_b200.so (cuobjdump -sass), written to profiles/ as the SASS evidence of a
round: which memory instructions (widths), atomics, votes/shuffles, FP64 ops each kernel is made of, plus registers/stack.
usage: python tools/sass_digest.py profiles/r02_sass_digest.txt"""
import collections, re, subprocess, sys, os

root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
lib = os.path.join(root, "mcut_b200", "lib", "libmcut_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", lib], capture_output=True, text=True).stdout
usage = {}
cur = None
for line in res.splitlines():
    m = re.match(r"\s*Function (\S+):", line)
    if m:
        cur = m.group(1)
    elif cur and "REG:" in line:
        usage[cur] = line.strip()
        cur = None
demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
kern, name = collections.OrderedDict(), None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        kern[name][m.group(1)] += 1
interesting = re.compile(r"^(LDG|STG|LDS|STS|LDL|STL|ATOM|ATOMS|ATOMG|RED|REDUX|VOTE|SHFL|MATCH|DFMA|DMUL|DADD|DSETP|DMNMX|FMNMX|BAR|MEMBAR|ACQBULK|LDGDEPBAR|DEPBAR|ERRBAR|CCTL|LDC|UTMA|SYNCS|WARPSYNC|NANOSLEEP|CALL|BSSY)")
out = [f"SASS digest of {os.path.relpath(lib, root)} (sm_100a), instruction counts per kernel (static), selected mnemonics", ""]
for k, c in kern.items():
    d = demangle(k)
    d = re.sub(r"\(anonymous namespace\)::", "", d)
    d = re.sub(r"\(.*", "", d)
    tot = sum(c.values())
    sel = sorted(((m, n) for m, n in c.items() if interesting.match(m)), key=lambda x: -x[1])
    out.append(f"{d}   [{tot} instructions; {usage.get(k, '')}]")
    line = "    "
    for m, n in sel:
        item = f"{m}={n}  "
        if len(line) + len(item) > 150:
            out.append(line.rstrip())
            line = "    "
        line += item
    out.append(line.rstrip())
    out.append("")
open(sys.argv[1], "w").write("\n".join(out))
print(len(kern), "kernels")
