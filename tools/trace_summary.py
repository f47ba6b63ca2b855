#!/usr/bin/env python
"""Summarise gpurun_out/trace_dump_{fused,plain}.json (tools/trace_dump.py): per-phase times, A/B CTA split, warp wait/compute cycles."""
import json, statistics as st, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for name in ('fused', 'plain'):
    p = os.path.join(ROOT, 'gpurun_out', 'trace_dump_%s.json' % name)
    if not os.path.exists(p):
        continue
    d = json.load(open(p))[-1]
    live = [i for i, v in enumerate(d['start']) if v > -1e6]
    if len(live) < len(d['start']):
        t0 = min(d['start'][i] for i in live)   # older dumps hold empty rows
        d = {k: [ (v[i] - t0 if k in ('start','a_done','published','bar_exit','end','b_start') else v[i]) for i in live] for k, v in d.items()}
    n = len(d['smid'])
    first = {}
    for i, s in enumerate(d['smid']):
        first.setdefault(s, []).append(i)
    A = [min(v, key=lambda i: d['a_done'][i]) for v in first.values()]
    B = [max(v, key=lambda i: d['a_done'][i]) for v in first.values() if len(v) > 1]
    med = lambda k, idx: st.median([d[k][i] for i in idx])
    print('%s: %d CTAs on %d SMs | start max %.2f | loop done A med %.2f B med %.2f max %.2f | published max %.2f | bar_exit min %.2f max %.2f | end min %.2f max %.2f'
          % (name, n, len(first), max(d['start']), med('a_done', A), med('a_done', B) if B else -1, max(d['a_done']), max(d['published']),
             min(d['bar_exit']), max(d['bar_exit']), min(d['end']), max(d['end'])))
    if any(v > 0 for v in d['b_start']):
        # instrumented build: slot 4 (b_start) = the CTA starts waiting for the totals, slot 6 = totals gathered and reduced
        g6 = [((w << 32) | c) for w, c in d['w0_wait_comp']]
        print('   wait starts: min %.2f max %.2f | gathered: min %.2f max %.2f' % (min(d['b_start']), max(d['b_start']), min(d.get('gathered', [0])), max(d.get('gathered', [0]))))
    if False:
        for lab, idx in (('A', A), ('B', B)):
            if idx:
                w = st.median([d['w0_wait_comp'][i][0] for i in idx]); c = st.median([d['w0_wait_comp'][i][1] for i in idx])
                ls = st.median([d['b_start'][i] for i in idx]); le = st.median([d['a_done'][i] for i in idx])
                print('   %s warp0: wait %d + compute %d cycles; loop from %.2f to %.2f us -> %.0f MHz' % (lab, w, c, ls, le, (w + c) / max(le - ls, 1e-9)))
