#!/usr/bin/env python
"""Per-step ncu counters of the step kernels for every bench workload -> profiles/kernel_counters.json.

    python profiles/capture_counters.py [--workloads c2,c3,c3_late,c4,c5] [--tag r02]      (GPU box, under gpurun)

For each workload one `ncu` pass over `python bench.py --workload K ...` (the very command whose numbers the bench line
reports) collects, per launch of the step kernels, DRAM bytes read + written, warp instructions executed and the
serialised duration; the launches of a step (one for the one-kernel step, two -- move + paint -- otherwise) are summed
and averaged over the captured steps.  ncu flushes the caches before every launch it measures, which is the state the
bench's own L2 flush leaves behind, so the DRAM figure is the per-step traffic of the timed region.  The file carries
the hash of the CUDA sources (bench.kernel_source_sha): bench.py reports `issue_frac`, `dram_gbs_measured` and `traffic`
from it and says `counters_current: false` when the sources have changed since.  The launch lists (CSV) are kept next to
it as profiles/<tag>_launches_<workload>.csv.  Numbers printed by bench.py under ncu are never bench values.
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

METRICS = 'gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum'
STEP_KERNELS = ('step_fused_kernel', 'move_fast_kernel', 'move_kernel', 'paint_kernel', 'paint_normal_kernel')
SETTINGS = {      # steps captured, warm-up steps skipped
    'c2': (12, 6), 'c2_normal': (8, 4), 'c3': (8, 4), 'c3_late': (8, 4), 'c4': (5, 3), 'c5': (6, 3),
}


def to_bytes(value, unit):
    scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    return float(value.replace(',', '')) * scale[unit]


def to_us(value, unit):
    scale = {'ns': 1e-3, 'us': 1.0, 'usecond': 1.0, 'ms': 1e3, 'msecond': 1e3, 'nsecond': 1e-3, 'second': 1e6}
    return float(value.replace(',', '')) * scale[unit]


def parse(path):
    """Launch list -> [(kernel short name, {metric: value})] in launch order."""
    with open(path, newline='') as f:
        lines = [ln for ln in f if not ln.startswith('==')]
    rows = list(csv.DictReader(lines))
    launches, by_id = [], {}
    for r in rows:
        key = r['ID']
        if key not in by_id:
            name = r['Kernel Name']
            short = next((k for k in STEP_KERNELS if k in name), None)
            by_id[key] = (short, {'name': name})
            launches.append(by_id[key])
        m, unit, val = r['Metric Name'], r['Metric Unit'], r['Metric Value']
        d = by_id[key][1]
        if m == 'gpu__time_duration.sum':
            d['us'] = to_us(val, unit)
        elif m.startswith('dram__bytes'):
            d[m] = to_bytes(val, unit)
        elif m == 'smsp__inst_executed.sum':
            d['insts'] = float(val.replace(',', ''))
    return launches


def capture(key, tag, out_dir):
    steps, warm = SETTINGS.get(key, (6, 3))
    log = os.path.join(out_dir, '%s_launches_%s.csv' % (tag, key))
    cmd = ['ncu', '--metrics', METRICS, '--clock-control', 'none', '--csv', '--log-file', log,
           '-k', 'regex:' + '|'.join(STEP_KERNELS),
           sys.executable, os.path.join(ROOT, 'bench.py'), '--workload', key, '--steps', str(steps), '--warmup', str(warm),
           '--no-extra', '--no-cpu-baseline', '--no-parity', '--e2e-sync']
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        raise RuntimeError('ncu failed for %s:\n%s' % (key, res.stdout[-2000:]))
    line = [ln for ln in res.stdout.splitlines() if ln.startswith('{')][-1]
    cfg = json.loads(line)['config']
    launches = [(k, d) for k, d in parse(log) if k]
    # the device-timed loop comes first in bench.py: warm-up steps, then the timed ones; everything after is the e2e leg
    per_step = 1 if launches[0][0] == 'step_fused_kernel' else 2
    first = warm * per_step
    rows = launches[first:first + steps * per_step]
    assert len(rows) == steps * per_step, (key, len(launches), first, steps, per_step)
    kernels = {}
    for k, d in rows:
        a = kernels.setdefault(k, {'launches': 0, 'us': 0.0, 'dram': 0.0, 'insts': 0.0})
        a['launches'] += 1
        a['us'] += d['us']
        a['dram'] += d['dram__bytes_read.sum'] + d['dram__bytes_write.sum']
        a['insts'] += d['insts']
    out = {
        'envs_per_gpu': cfg['envs_per_gpu'], 'steps_captured': steps, 'launches_per_step': per_step,
        'dram_bytes_per_step': sum(a['dram'] for a in kernels.values()) / steps,
        'warp_insts_per_step': sum(a['insts'] for a in kernels.values()) / steps,
        'serialized_us_per_step': sum(a['us'] for a in kernels.values()) / steps,
        'kernels': {k: {'us': a['us'] / a['launches'], 'dram_bytes': a['dram'] / a['launches'], 'warp_insts': a['insts'] / a['launches'],
                        'share_of_step': a['us'] / sum(b['us'] for b in kernels.values())} for k, a in kernels.items()},
        'launch_list': os.path.relpath(log, ROOT),
    }
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--workloads', default='c2,c3,c3_late,c4,c5')
    ap.add_argument('--tag', default='r02')
    ap.add_argument('--out-dir', default=os.path.join(ROOT, 'gpurun_out'), help='where the launch lists and the JSON go (copy them into profiles/)')
    args = ap.parse_args()
    os.makedirs(args.out_dir, exist_ok=True)
    doc = {'source_sha': bench.kernel_source_sha(), 'metrics': METRICS, 'ncu': 'one pass per workload, --clock-control none, caches flushed per launch (ncu default)',
           'workloads': {}}
    for key in args.workloads.split(','):
        doc['workloads'][key] = capture(key, args.tag, args.out_dir)
        w = doc['workloads'][key]
        print('%-8s %d launch(es) per step, serialised %.1f us, DRAM %.2f MB, %.2f M warp instructions per step' % (
            key, w['launches_per_step'], w['serialized_us_per_step'], w['dram_bytes_per_step'] / 1e6, w['warp_insts_per_step'] / 1e6), flush=True)
    with open(os.path.join(args.out_dir, 'kernel_counters.json'), 'w') as f:
        json.dump(doc, f, indent=1, sort_keys=True)
    print('wrote', os.path.join(args.out_dir, 'kernel_counters.json'), 'for sources', doc['source_sha'])


if __name__ == '__main__':
    main()
