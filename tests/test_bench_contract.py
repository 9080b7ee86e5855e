"""The bench.py JSON contract, checked on the committed bench lines (profiles/bench_r0*_*.json): every
key the driver and the judge read must be present with the right type, and the derived figures must be
consistent (roofline.frac = achieved / peak, achieved = algorithmic bytes x realisations / kernel time)."""
import glob
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LINES = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'bench_r0*_*.json')))


def _load(path):
    for line in open(path):
        if line.startswith('{'):
            return json.loads(line)
    raise AssertionError(f'no JSON line in {path}')


@pytest.mark.parametrize('path', LINES, ids=[os.path.basename(p) for p in LINES])
def test_bench_line_has_the_contract_keys(path):
    d = _load(path)
    for key, typ in (('metric', str), ('value', float), ('unit', str), ('n_gpus', int), ('steps', int),
                     ('warmup', int), ('ms_per_step', float), ('higher_is_better', bool), ('scaling', str),
                     ('dtype', str), ('data', str), ('config', dict), ('e2e', dict)):
        assert key in d and isinstance(d[key], typ), key
    assert 'vs_baseline' in d and d['vs_baseline'] is None          # BASELINE.md publishes no number
    assert d['higher_is_better'] is True and d['scaling'] == 'weak' and 'workload' in d['config']
    assert set(d['e2e']) >= {'value', 'unit', 'h2d_bytes_per_step', 'd2h_bytes_per_step'}
    cb = d.get('cpu_baseline')          # N = 1 only, and absent from the --no-cpu side runs (c4, c5)
    assert cb is None or (set(cb) >= {'value', 'unit', 'cores', 'kind', 'sample'} and cb['kind'] == 'port')
    if d.get('impl') == 'reference':
        assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['value'] == d['value'] and d['gpu_launches'] == 0
        return
    assert d['warmup'] >= 3 and d['gpu_launches'] > 0
    r = d['roofline']
    assert set(r) >= {'bound', 'achieved', 'peak', 'unit', 'frac', 'traffic'} and r['bound'] == 'hbm' and r['unit'] == 'GB/s'
    assert r['frac'] == pytest.approx(r['achieved'] / r['peak'], rel=1e-9)
    want = r['algorithmic_bytes_per_realization'] * r['realizations_per_launch'] / (r['kernel_ms'] * 1e-3) / 1e9
    assert r['achieved'] == pytest.approx(want, rel=1e-9)
    c = d['clocks']
    assert set(c) >= {'sm_mhz', 'sm_max_mhz', 'reasons'}
    assert not any('thermal' in x or 'hw_slowdown' in x for x in c['reasons'])
    assert d['e2e']['h2d_bytes_per_step'] > 0 and d['e2e']['d2h_bytes_per_step'] > 0
    # whole-job throughput is consistent with the step time
    n_real = r['realizations_per_launch'] * d['n_gpus']
    assert d['value'] == pytest.approx(n_real / (d['ms_per_step'] * 1e-3), rel=0.02)


R02 = sorted(glob.glob(os.path.join(ROOT, 'profiles', 'bench_r02*_c2_*gpu.json')))


@pytest.mark.parametrize('path', R02, ids=[os.path.basename(p) for p in R02])
def test_round2_lines_carry_the_evidence_the_verdict_asked_for(path):
    """e2e through Fast.run() comes out below the device-timed value, a step is >= 100 ms and one launch,
    every configuration has a per_workload entry, and N > 1 lines carry the sharded bit-identity check."""
    d = _load(path)
    assert d['e2e']['value'] <= d['value'] and 'Fast(p).run()' in d['e2e']['what']
    assert d['ms_per_step'] >= 100.0 and d['gpu_launches'] == d['steps']
    assert d['check']['e2e_result_ok'] is True
    pw = d['per_workload']
    if d['n_gpus'] == 1:
        assert set(pw) >= {'c1_temporal', 'c1prime', 'c2', 'c2_device_fast', 'c3_sweep', 'c4', 'c5', 'c5_strong'}
        assert d['check']['sharded_bit_identical'] is None
        assert pw['c1prime']['value'] > 5 * pw['c1prime']['direct_dft_value']
        assert pw['c2_device_fast']['value'] > 1.05 * pw['c2']['value']
    else:
        assert set(pw) >= {'c3_sweep', 'c5_strong'} and d['check']['sharded_bit_identical'] is True
        assert pw['c5_strong']['n_reduced'] == 1000000 and pw['c5_strong']['n_gpus'] == d['n_gpus']
    for name, e in pw.items():
        if 'roofline_frac' in e:
            want = e['algorithmic_bytes_per_realization'] * e['value'] / e['n_gpus'] / 1e9 / d['roofline']['peak']
            assert e['roofline_frac'] == pytest.approx(want, rel=1e-9), name
    sec = d['roofline'].get('secondary')
    if sec is not None:
        assert sec['bound'] == 'issue' and sec['frac'] == pytest.approx(sec['achieved_ginst_s'] / sec['peak_ginst_s'])
        assert d['roofline']['traffic'] is not None


def test_headline_line_is_the_c2_workload_on_one_gpu():
    d = _load(os.path.join(ROOT, 'profiles', 'bench_r01_c2.json'))
    assert d['n_gpus'] == 1 and 'C2' in d['config']['workload'] and d['dtype'] == 'f32'
    assert d['roofline']['algorithmic_bytes_per_realization'] == 8 * 256 * 256 + 4
    assert d['k5_link_metrics'] and d['comparator'] and d['k1_psd_build']
    assert d['cpu_baseline']['cores'] == 1 and d['cpu_baseline']['value'] > 0
