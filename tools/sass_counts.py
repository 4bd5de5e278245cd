"""SASS evidence (runs here, no GPU): per kernel of libswat_b200.so, how many tcgen05 / TMEM / TMA instructions it holds.
    python tools/sass_counts.py > profiles/r02_sass_counts.md"""
import collections, re, subprocess, sys
so = sys.argv[1] if len(sys.argv) > 1 else "swat_b200/libswat_b200.so"
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|LDTM(?:\.x\d+)?|UTMALDG\.2D(?:\.2CTA)?|UTCATOM\w*|SYNCS\.ARRIVE\.TRANS64\w*(?:\.\w+)*|SYNCS\.PHASECHK\.TRANS64\.TRYWAIT|F2FP\.BF16\.F32\.PACK_AB|FENCE\.VIEW\.ASYNC\.S|HMMA\S*|RED\.E\.ADD\S*|MEMBAR\.\S+)")
kern, counts, order = None, collections.defaultdict(collections.Counter), []
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(anonymous namespace\)::|swat::|\(CUtensorMap_st.*|\(swat::.*|\(.*", "", kern).replace("void ", "")
        order.append(kern)
        continue
    if kern:
        for op in pat.findall(line):
            counts[kern][re.sub(r"\.x\d+", "", op)] += 1
groups = collections.OrderedDict()
for k in order:
    base = re.sub(r"<.*", "", k)
    groups.setdefault(base, []).append(k)
print(f"# SASS mnemonics per kernel of `{so}` (cuobjdump -sass, sm_100a)\n")
print("`UTCHMMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, `UTCBAR` = tcgen05.commit, `SYNCS.*` = mbarrier, "
      "`F2FP.BF16.F32.PACK_AB` = the fp32 -> bf16 converter warps, `FENCE.VIEW.ASYNC` = fence.proxy.async.  No `HMMA` (legacy mma.sync) anywhere.\n")
print("| kernel | instantiations | UTCHMMA | LDTM | UTMALDG | UTCBAR | F2FP pack | mbarrier arrive / try_wait | HMMA | MEMBAR |")
print("|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|")
for base, ks in groups.items():
    tot = collections.Counter()
    for k in ks:
        tot.update(counts[k])
    g = lambda p: sum(v for kk, v in tot.items() if kk.startswith(p))
    print(f"| `{base}` | {len(ks)} | {g('UTCHMMA')} | {g('LDTM')} | {g('UTMALDG')} | {g('UTCBAR')} | {g('F2FP')} | "
          f"{g('SYNCS.ARRIVE')} / {g('SYNCS.PHASECHK')} | {g('HMMA')} | {g('MEMBAR')} |")
print("\nPer instantiation of the scan kernel `scan_tc_kernel<ctas, reduce, partitioned, dense, fp32-bank>`:\n")
print("| instantiation | UTCHMMA | LDTM | UTMALDG | F2FP pack |")
print("|---|---:|---:|---:|---:|")
for k in groups.get("scan_tc_kernel", []):
    c = counts[k]
    g = lambda p: sum(v for kk, v in c.items() if kk.startswith(p))
    print(f"| `{k[k.index('<'):]}` | {g('UTCHMMA')} | {g('LDTM')} | {g('UTMALDG')} | {g('F2FP')} |")
