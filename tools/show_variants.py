import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l)
    if 'shape' in d:
        print('%-38s %-34s theta %.3f ms (%.3f)  occ-in %.3f ms (%.3f)  occ-kernel %.3f  3xtf32 %s' % (
            d['shape'], d.get('tune', ''), d['theta_ms'], d['executed_frac_theta'], d['occ_input_ms'],
            d['executed_frac_occ_input'], d['occupation_kernel_ms'],
            '%.3f ms' % d['tf32_theta_ms'] if d.get('tf32_theta_ms') else '-'))
