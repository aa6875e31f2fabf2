#!/usr/bin/env python
"""Join an ncu SASS source page (csv) with nvdisasm -g line info: executed warp-instructions and stall samples per
source line.   python tools/sass_lines.py <disasm> <ncu_sass.csv> <kernel mangled substring> <warps> [top]"""
import collections, csv, re, sys

disasm, prof_csv, key, warps = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
txt = open(disasm).read().splitlines()
start = next(i for i, l in enumerate(txt) if l.strip().startswith('.text.') and key in l and l.strip().endswith(':'))
seq, cur = [], None
for l in txt[start + 1:]:
    if l.strip().startswith('.text.') and l.strip().endswith(':'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s*/\*([0-9a-f]{4,})\*/\s+(.*?);', l)
    if m:
        seq.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(prof_csv)))
ends = [i for i, r in enumerate(rows) if r and r[0] == 'Kernel Name']
if len(ends) > 1:  # several launches in the report: the first one
    rows = rows[:ends[1]]
hdr = rows[1]
ia, ie, iss = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Warp Stall Sampling (All Samples)')
prof = [(r[ia].strip(), int(r[ie]), int(r[iss]) if r[iss].isdigit() else 0) for r in rows[2:] if len(r) >= len(hdr)]
assert len(seq) == len(prof), (len(seq), len(prof))
by, st, ops = collections.Counter(), collections.Counter(), collections.Counter()
for (a, ln), (b, e, s) in zip(seq, prof):
    by[ln] += e
    st[ln] += s
    o = b.split()
    o = (o[1] if o[0].startswith('@') else o[0]).split('.')[0]
    ops[o] += e
tot = sum(by.values())
print(f'total {tot / warps:.1f} warp-instr per warp; stall samples {sum(st.values())}')
print('  '.join(f'{o}:{n / warps:.0f}' for o, n in ops.most_common(24)))
src = {}
for (f, ln), e in by.most_common(top):
    if f not in src:
        try:
            src[f] = open(__import__('os').environ.get('KZN_SRC_DIR', '/root/repo/kazan_b200/csrc/') + f).read().splitlines()
        except OSError:
            src[f] = None
    text = src[f][ln - 1].strip()[:95] if src[f] and ln <= len(src[f]) else ''
    print(f'{f}:{ln:<5} {e / warps:7.1f}  stall {st[(f, ln)]:5}  {text}')
