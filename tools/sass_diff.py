#!/usr/bin/env python
"""Compare the SASS of every kernel of two builds (directories of `cuobjdump -sass` dumps, one file per object).

Used to show that a change which only ADDS template variants / parameters leaves the measured default kernels alone:
instructions are compared after masking what legitimately moves (constant-bank offsets of kernel parameters, branch
targets, immediates); a kernel whose template gained trailing boolean parameters is matched to its all-`false` instantiation.

    python tools/sass_diff.py /tmp/orig/sass /tmp/new/sass
"""
import difflib
import os
import re
import sys


def funcs(path):
    out, cur = {}, None
    for line in open(path):
        m = re.search(r'Function : (\S+)', line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        if cur and re.match(r'\s+/\*[0-9a-f]{4,6}\*/', line):
            ins = re.sub(r'/\*[0-9a-fx ]+\*/', '', line).strip()
            out[cur].append(re.sub(r'0x[0-9a-f]+', 'H', ins))
    return out


def key(name):                       # kernel name + template arguments (the parameter list may have grown), and a template
    head = name.split('EEv')[0]      # that GAINED trailing boolean parameters matches its all-`false` instantiation
    return re.sub(r'(Lb0E)+$', '', head)


def main(a, b):
    worst = 0
    for f in sorted(os.listdir(a)):
        if not f.endswith('.sass') or not os.path.exists(os.path.join(b, f)):
            continue
        old, new = funcs(os.path.join(a, f)), funcs(os.path.join(b, f))
        for n, v in old.items():
            cands = [n] if n in new else [c for c in new if key(c) == key(n) and c.split('EEv')[0] != n.split('EEv')[0] or key(c) == n.split('EEv')[0]]
            if not cands:
                print(f'{f}: {n[:70]}: no counterpart')
                worst = max(worst, 10 ** 6)
                continue
            w = new[cands[0]]
            ops = [o for o in difflib.SequenceMatcher(None, v, w, autojunk=False).get_opcodes() if o[0] != 'equal']
            changed = sum(max(o[2] - o[1], o[4] - o[3]) for o in ops)
            # register renames show up as 1:1 replacements of the same opcode
            real = sum(max(o[2] - o[1], o[4] - o[3]) for o in ops
                       if not (o[0] == 'replace' and o[2] - o[1] == o[4] - o[3] and
                               all(x.split()[0:1] == y.split()[0:1] or (x.startswith('@') and y.startswith('@'))
                                   for x, y in zip(v[o[1]:o[2]], w[o[3]:o[4]]))))
            worst = max(worst, real)
            tag = 'identical' if not ops else f'{changed} changed ({real} beyond same-opcode operand renames)'
            print(f'{f}: {n[4:64]:60s} {len(v):5d} -> {len(w):5d}  {tag}')
    print('max structural difference in any default kernel:', worst, 'instructions')


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2])
