import json
d=json.load(open('gpurun_out/r3.out')); print('RESULT', d['k3_units'], d['batches'], d['memory_cap_units'], d['gpu_ms'], d['parity']['max_rel_err_above_1e-6_of_max'])
