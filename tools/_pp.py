import sys, json
for l in sys.stdin:
    if l.startswith('{') and 'ms_per_step' in l:
        d = json.loads(l); print(round(d['ms_per_step'], 4), d['steps'], d['rejected'], d['rhs'])
