#!/usr/bin/env python
"""pretty-print the JSON line of a bench.py log (measurement aid): python tools/show_bench.py gpurun_out/x.log"""
import json, sys
def show(x, ind=0, maxlen=150):
    for k, v in x.items():
        if isinstance(v, dict):
            print(' ' * ind + k + ':'); show(v, ind + 2)
        elif isinstance(v, list) and v and isinstance(v[0], dict):
            for i, e in enumerate(v):
                print(' ' * ind + f'{k}[{i}]:'); show(e, ind + 2)
        else:
            print(' ' * ind + f'{k}: {str(v)[:maxlen]}')
for line in open(sys.argv[1]):
    if line.startswith('{"metric"') or line.startswith('{"impl"'):
        show(json.loads(line))
