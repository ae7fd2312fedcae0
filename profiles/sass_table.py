"""Table of SASS mnemonics per kernel: python profiles/sass_table.py > table.md (reads dge_b200/_build/*.o)."""
import glob, os, re, subprocess
rows = []


def kernel_name(sig):
    """The demangled signature without its parameter list (template arguments like <(bool)1> stay)."""
    depth = 0
    for i, ch in enumerate(sig):
        if ch == "<": depth += 1
        elif ch == ">": depth -= 1
        elif ch == "(" and depth == 0: return sig[:i].strip()
    return sig.strip()


for o in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "dge_b200", "_build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
    demangled = subprocess.run(["cu++filt"], input=sass, capture_output=True, text=True).stdout or sass
    cur, cnt = None, None
    for line in demangled.splitlines():
        m = re.search(r"Function : (.+)", line)
        if m:
            if cur: rows.append((os.path.basename(o), cur, cnt))
            cur, cnt = kernel_name(m.group(1)), dict.fromkeys(("n", "UBLKCP", "SYNCS", "MUFU.EX2", "MUFU.RCP", " REDG", "SHFL", "STL", "LDL"), 0)
            continue
        if cur and re.search(r"/\*[0-9a-f]{4}\*/", line):
            cnt["n"] += 1
            for k in cnt:
                if k != "n" and k in line: cnt[k] += 1
    if cur: rows.append((os.path.basename(o), cur, cnt))
print("| object | kernel | instructions | UBLKCP (TMA bulk copy) | SYNCS (mbarrier) | MUFU.EX2 | MUFU.RCP | REDG | SHFL | STL/LDL (spills) |\n|---|---|---:|---:|---:|---:|---:|---:|---:|---:|")
for o, k, c in rows:
    print(f"| {o} | `{k}` | {c['n']} | {c['UBLKCP']} | {c['SYNCS']} | {c['MUFU.EX2']} | {c['MUFU.RCP']} | {c[' REDG']} | {c['SHFL']} | {c['STL'] + c['LDL']} |")
