"""CPU checks of bench.py's reference arm (`--impl reference`): the JSON contract of the line and the accounting of the bounded
sample (microbatches + one AdamW step, reported as the rate of a whole device batch).  The oracle network is swapped for the
tiny configuration so that the test runs in seconds; the arithmetic of the line does not depend on the model size."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_reference_arm_line(monkeypatch, capsys):
    import oracle.unet
    monkeypatch.setattr(oracle.unet, 'SD2_BASE_UNET_CONFIG', oracle.unet.TINY_UNET_CONFIG)
    import bench
    rate, t_mb, t_opt = bench.oracle_cpu_step_rate(32, 2, 2, 1, 2, 256)
    assert t_mb > 0 and t_opt > 0
    assert abs(rate - 256 / (128 * t_mb + t_opt)) < 1e-9 * rate  # 128 microbatches of 2 images + one optimizer step
    monkeypatch.setenv('RANK', '0')
    args = argparse.Namespace(latent=32, steps=1, warmup=1, gpus=1, batch=256)
    bench.run_reference(args)
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == bench.METRIC and line['unit'] == 'images/s'
    assert line['higher_is_better'] is True and line['vs_baseline'] is None and line['n_gpus'] == 1 and line['steps'] == 1
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['value'] == line['value'] and cb['cores'] == (os.cpu_count() or 1) and 'AdamW' in cb['sample']
    assert line['config']['per_gpu_microbatch'] == 256 and 'AdamW' in line['config']['workload']
    # ranks other than 0 print nothing and return (the driver launches the arm under torchrun for N > 1)
    monkeypatch.setenv('RANK', '1')
    bench.run_reference(args)
    assert capsys.readouterr().out == ''
