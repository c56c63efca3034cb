import sys, json
d = json.loads(sys.stdin.read()); print('RESULT', d['k3_units'], d['batches'], d['gpu_ms'], d['parity']['max_rel_err_above_1e-6_of_max'])
