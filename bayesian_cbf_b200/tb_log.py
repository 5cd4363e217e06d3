"""Run logs in the reference's TensorBoard format, so that its `*_vis` / `*_plot` scripts can read runs made
with this package.

The reference logs scalars with SummaryWriter.add_scalar and whole arrays as DT_FLOAT TensorProto summaries
(bayes_cbf/misc.py:320-359: make_tensor_summary / add_tensors), and reads both back with an EventFileLoader
(misc.py:342-359: stream_tensorboard_scalars / load_tensorboard_scalars; callers pendulum.py:496,1219,1409,
unicycle_move_to_pose.py, trigger_interval.py:104).  This module writes the same records with tensorboard's own
EventFileWriter (no torch.utils.tensorboard / TensorFlow needed) and keeps the reference's names:

    logger = TBLogger(['unicycle', 'safe'], runs_dir='data/runs')      # or NoLogger()
    logger.add_scalars('opt', dict(loss=0.3), t)
    logger.add_tensors('traj', dict(x=x, u=u, dx=dx), t)
    by_tag = load_tensorboard_scalars(events_file)                      # {tag: [(t, value), ...]}
"""
import glob
import os
import os.path as osp
import time
from abc import ABC, abstractmethod

import numpy as np


def _protos():
    from tensorboard.compat.proto import event_pb2, summary_pb2, tensor_pb2, tensor_shape_pb2
    return event_pb2, summary_pb2, tensor_pb2, tensor_shape_pb2


def _as_array(v):
    if hasattr(v, 'detach'):
        v = v.detach().cpu().numpy()
    return np.asarray(v, dtype=np.float64)


def make_tensor_summary(name, nparray):
    """Summary holding one array as a DT_FLOAT TensorProto with explicit float_val and shape (misc.py:320-326)."""
    _, summary_pb2, tensor_pb2, tensor_shape_pb2 = _protos()
    a = _as_array(nparray)
    shape = tensor_shape_pb2.TensorShapeProto(dim=[tensor_shape_pb2.TensorShapeProto.Dim(size=int(s)) for s in a.shape])
    tensor = tensor_pb2.TensorProto(dtype='DT_FLOAT', float_val=a.reshape(-1).tolist(), tensor_shape=shape)
    return summary_pb2.Summary(value=[summary_pb2.Summary.Value(tag=name, tensor=tensor)])


def make_scalar_summary(name, value):
    _, summary_pb2, _, _ = _protos()
    return summary_pb2.Summary(value=[summary_pb2.Summary.Value(tag=name, simple_value=float(value))])


class EventWriter:
    """Minimal stand-in for the SummaryWriter the reference passes around: an events.out.tfevents.* file in `logdir`."""

    def __init__(self, logdir):
        from tensorboard.summary.writer.event_file_writer import EventFileWriter
        os.makedirs(logdir, exist_ok=True)
        self.logdir = logdir
        self._writer = EventFileWriter(logdir)

    def add_summary(self, summary, step):
        event_pb2 = _protos()[0]
        self._writer.add_event(event_pb2.Event(wall_time=time.time(), step=int(step), summary=summary))

    def add_scalar(self, tag, value, step):
        self.add_summary(make_scalar_summary(tag, value), step)

    def flush(self):
        self._writer.flush()

    def close(self):
        self._writer.close()

    def event_files(self):
        return sorted(glob.glob(osp.join(self.logdir, 'events.out.tfevents.*')))


def add_tensors(summary_writer, tag, var_dict, t):
    """One tensor summary per entry, tag "<tag>/<key>" (misc.py:329-335)."""
    for k, v in var_dict.items():
        summary_writer.add_summary(make_tensor_summary("/".join((tag, k)), v), t)


def stream_tensorboard_scalars(event_file):
    """(step, tag, value) for every summary of an event file; value is a float or an array of the logged shape
    (misc.py:342-352).  Newer tensorboard loaders deliver scalars as rank-0 tensors: both forms are handled."""
    from tensorboard.backend.event_processing import event_file_loader
    for event in event_file_loader.EventFileLoader(event_file).Load():
        if event.summary is None or not len(event.summary.value):
            continue
        val = event.summary.value[0]
        if val.HasField('tensor') and (len(val.tensor.float_val) or len(val.tensor.tensor_shape.dim)):
            value = np.array(val.tensor.float_val).reshape([d.size for d in val.tensor.tensor_shape.dim])
            if value.ndim == 0:
                value = float(value)
        else:
            value = val.simple_value
        yield event.step, val.tag, value


def load_tensorboard_scalars(event_file):
    """{tag: [(step, value), ...]} in file order (misc.py:355-359)."""
    by_tag = dict()
    for t, tag, value in stream_tensorboard_scalars(event_file):
        by_tag.setdefault(tag, []).append((t, value))
    return by_tag


class Logger(ABC):
    @property
    @abstractmethod
    def experiment_logs_dir(self):
        return "/tmp"

    @abstractmethod
    def add_scalars(self, tag, var_dict, t):
        pass

    @abstractmethod
    def add_tensors(self, tag, var_dict, t):
        pass


class NoLogger(Logger):
    @property
    def experiment_logs_dir(self):
        return "/tmp"

    def add_scalars(self, tag, var_dict, t):
        pass

    def add_tensors(self, tag, var_dict, t):
        pass


class TBLogger(Logger):
    """Reference misc.py:386-405: one run directory `<runs_dir>/<tags joined by _>_<version>`."""

    def __init__(self, exp_tags, runs_dir='data/runs', version='b200'):
        self.exp_tags = list(exp_tags)
        self.runs_dir = runs_dir
        self.exp_dir = osp.join(runs_dir, '_'.join(self.exp_tags + [version]))
        self.summary_writer = EventWriter(self.exp_dir)

    @property
    def experiment_logs_dir(self):
        return self.exp_dir

    def add_scalars(self, tag, var_dict, t):
        for k, v in var_dict.items():
            self.summary_writer.add_scalar("/".join((tag, k)), float(_as_array(v)), t)

    def add_tensors(self, tag, var_dict, t):
        add_tensors(self.summary_writer, tag, var_dict, t)

    def flush(self):
        self.summary_writer.flush()

    def close(self):
        self.summary_writer.close()


def ensuredirs(fpath):
    fdir = osp.dirname(fpath)
    if fdir and not osp.exists(fdir):
        os.makedirs(fdir)
    return fpath
