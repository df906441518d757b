import json, sys
sys.path.insert(0, '.')
import torch, bench
for b in (32, 256, 1024):
    print(json.dumps(bench.measure_next_rows(torch, batch=b)['iqn_loss']))
