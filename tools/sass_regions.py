#!/usr/bin/env python
"""Stall samples and executed instructions per source line and per code region of raster_kernel.
   python tools/sass_regions.py <disasm> <ncu_source.csv> <kernel mangled substring> <warps> <source file>"""
import re, csv, collections, sys
disasm, prof_csv, key, warps, srcfile = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4]), sys.argv[5]
txt = open(disasm).read().splitlines()
start = next(i for i, l in enumerate(txt) if l.strip().startswith('.text.') and key in l and l.strip().endswith(':'))
seq = []; cur = None
for l in txt[start + 1:]:
    if l.strip().startswith('.text.') and l.strip().endswith(':'): break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m: seq.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(prof_csv)))
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
if len(ends) > 1: rows = rows[:ends[1]]
hdr = rows[1]
ia, ie, iss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
prof = [(r[ia].strip(), int(r[ie]), int(r[iss]) if r[iss].isdigit() else 0) for r in rows[2:] if len(r) >= len(hdr)]
assert len(seq) == len(prof), (len(seq), len(prof))
st = collections.Counter(); ins = collections.Counter(); opst = collections.Counter()
for (a, ln), (b, e, s) in zip(seq, prof):
    st[ln] += s; ins[ln] += e
    o = b.split(); o = (o[1] if o[0].startswith('@') else o[0])
    opst[o.split('.')[0]] += s
tot = sum(st.values()); toti = sum(ins.values())
print('stall samples by opcode:', opst.most_common(14))
src = open(srcfile).read().splitlines()
# regions: marked in the source by lines containing "// @region <name>"
marks = [(i + 1, l.split('@region')[1].strip()) for i, l in enumerate(src) if '@region' in l]
def region(ln):
    if ln is None: return 'other'
    f, l = ln
    if not srcfile.endswith(f): return 'intrinsics'
    r = 'helpers'
    for at, name in marks:
        if l >= at: r = name
    return r
for (f, ln), s in st.most_common(36):
    t = src[ln - 1].strip()[:100] if srcfile.endswith(f) else f
    print(f'{ln:5} stall {s:5} {100*s/tot:4.1f}%  instr {ins[(f,ln)]/warps:6.1f}  {t}')
rs = collections.Counter(); ri = collections.Counter()
for ln, s in st.items(): rs[region(ln)] += s
for ln, e in ins.items(): ri[region(ln)] += e
print(f'total {toti/warps:.0f} warp-instr per warp')
for k in sorted(rs, key=lambda k: -rs[k]): print(f'{k:14} stall {100*rs[k]/tot:5.1f}%  instr {100*ri[k]/toti:5.1f}%')
