"""Run logs in the reference's TensorBoard tensor-summary format (bayes_cbf/misc.py:320-359): what TBLogger writes is read
back by load_tensorboard_scalars, and the raw records have exactly the fields the reference's reader looks at
(tag, tensor.float_val, tensor.tensor_shape.dim, step)."""
import numpy as np
import pytest
import torch

pytest.importorskip('tensorboard')

from bayesian_cbf_b200 import tb_log


def test_tensor_and_scalar_round_trip(tmp_path):
    logger = tb_log.TBLogger(['unicycle', 'safe'], runs_dir=str(tmp_path))
    assert logger.experiment_logs_dir.endswith('unicycle_safe_b200')
    x = np.arange(12, dtype=np.float64).reshape(3, 4) / 7.0
    u = torch.linspace(-1, 1, 5, dtype=torch.float64)
    for t in range(3):
        logger.add_tensors('traj', dict(x=x + t, u=u * t), t)
        logger.add_scalars('opt', dict(loss=0.5 ** t), t)
    logger.close()
    files = logger.summary_writer.event_files()
    assert len(files) == 1
    data = tb_log.load_tensorboard_scalars(files[0])
    assert set(data) == {'traj/x', 'traj/u', 'opt/loss'}
    for t in range(3):
        step, val = data['traj/x'][t]
        assert step == t and val.shape == (3, 4)
        assert np.allclose(val, (x + t).astype(np.float32))            # DT_FLOAT: float32 on disk, as in the reference
        assert np.allclose(data['traj/u'][t][1], (u * t).numpy().astype(np.float32))
        assert data['opt/loss'][t] == (t, 0.5 ** t)


def test_raw_records_have_the_fields_the_reference_reads(tmp_path):
    from tensorboard.backend.event_processing import event_file_loader
    w = tb_log.EventWriter(str(tmp_path / 'run'))
    tb_log.add_tensors(w, 'train', dict(Xtrain=np.ones((2, 3))), 7)
    w.close()
    events = [e for e in event_file_loader.EventFileLoader(w.event_files()[0]).Load() if len(e.summary.value)]
    assert len(events) == 1 and events[0].step == 7
    val = events[0].summary.value[0]
    assert val.tag == 'train/Xtrain'
    assert list(val.tensor.float_val) == [1.0] * 6 and [d.size for d in val.tensor.tensor_shape.dim] == [2, 3]


def test_nologger_is_silent():
    lg = tb_log.NoLogger()
    lg.add_scalars('a', dict(b=1.0), 0)
    lg.add_tensors('a', dict(b=np.zeros(2)), 0)
    assert lg.experiment_logs_dir == '/tmp'


@pytest.mark.skipif(not __import__('os').path.isdir('/root/reference/bayes_cbf'), reason="needs the reference checkout")
def test_reference_reader_reads_our_logs(tmp_path):
    """SURVEY 8f-4 from the consumer side: the reference's OWN `misc.load_tensorboard_scalars` (misc.py:342-359, imported
    unmodified in a subprocess; matplotlib stubbed) reads a run written by `tb_log.TBLogger` and recovers every array."""
    import subprocess
    import sys
    logger = tb_log.TBLogger(['pendulum', 'speed'], runs_dir=str(tmp_path))
    x = np.arange(6, dtype=np.float64).reshape(2, 3) / 4.0
    for t in range(4):
        logger.add_tensors('traj', dict(x=x * (t + 1), u=np.array([0.25 * t])), t)
        logger.add_scalars('opt', dict(loss=1.0 / (t + 1)), t)
    logger.close()
    path = logger.summary_writer.event_files()[0]
    code = (
        "import sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "from oracle import gpytorch_shim; gpytorch_shim.install()\n"
        "from bayes_cbf.misc import load_tensorboard_scalars\n"
        "d = load_tensorboard_scalars(%r)\n"
        "assert set(d) == {'traj/x', 'traj/u', 'opt/loss'}, set(d)\n"
        "x = np.arange(6, dtype=np.float64).reshape(2, 3) / 4.0\n"
        "for t in range(4):\n"
        "    step, val = d['traj/x'][t]\n"
        "    assert step == t and np.asarray(val).shape == (2, 3) and np.allclose(val, x * (t + 1))\n"
        "    assert np.allclose(d['traj/u'][t][1], [0.25 * t])\n"
        "    assert abs(float(np.asarray(d['opt/loss'][t][1]).reshape(-1)[0]) - 1.0 / (t + 1)) < 1e-6\n"
        "print('REF_READER_OK')\n" % (__import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))), path))
    res = subprocess.run([sys.executable, '-c', code], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and 'REF_READER_OK' in res.stdout, res.stdout[-1500:] + res.stderr[-3000:]
