"""marius_b200.host -- the C++/libtorch adapters (marius_b200/csrc/host/*.cpp, pybind module lib/_host) under the names the
reference's Python bindings use (src/cpp/python_bindings: marius.storage / marius.nn / marius.data), so a script written against

    from marius.nn.decoders.edge import DistMult ; from marius.data import Batch ; from marius.nn import Model, SoftmaxCrossEntropy

reads the same with `from marius_b200.host import nn, data, storage`.  Everything here runs the sm_100a kernels through the
C ABI; there is no CPU fallback (constructing device objects without CUDA raises MariusRuntimeException)."""
from __future__ import annotations

import importlib.util
import os
import sysconfig
from types import SimpleNamespace

from . import _lib  # noqa: F401  loads libmarius_b200.so (RTLD_GLOBAL) before the extension that links against it

_HERE = os.path.dirname(os.path.abspath(__file__))
_EXT = os.path.join(_HERE, "lib", "_host" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
if not os.path.exists(_EXT):
    raise ImportError(f"{_EXT} not found: build it with `python -m marius_b200.build`")

import torch  # noqa: E402,F401  (libtorch must be loaded first)

_spec = importlib.util.spec_from_file_location("_host", _EXT)
_host = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_host)

Storage, InMemory, PartitionBuffer, PartitionBufferStorage = _host.Storage, _host.InMemory, _host.PartitionBuffer, _host.PartitionBufferStorage
EdgeDecoder, DistMult, ComplEx = _host.EdgeDecoder, _host.DistMult, _host.ComplEx
Batch, Model, LossFunction, SoftmaxCrossEntropy = _host.Batch, _host.Model, _host.LossFunction, _host.SoftmaxCrossEntropy
node_corrupt_forward, only_pos_forward = _host.node_corrupt_forward, _host.only_pos_forward
LinkPredictionReporter, Hitsk, MeanRank, MeanReciprocalRank = _host.LinkPredictionReporter, _host.Hitsk, _host.MeanRank, _host.MeanReciprocalRank
ComputeWorkerGPU = _host.ComputeWorkerGPU
MariusRuntimeException = _host.MariusRuntimeException
set_default_precision, default_precision = _host.set_default_precision, _host.default_precision

storage = SimpleNamespace(Storage=Storage, InMemory=InMemory, PartitionBuffer=PartitionBuffer, PartitionBufferStorage=PartitionBufferStorage)
data = SimpleNamespace(Batch=Batch)
pipeline = SimpleNamespace(ComputeWorkerGPU=ComputeWorkerGPU)
report = SimpleNamespace(LinkPredictionReporter=LinkPredictionReporter, Hitsk=Hitsk, MeanRank=MeanRank, MeanReciprocalRank=MeanReciprocalRank)
nn = SimpleNamespace(Model=Model, SoftmaxCrossEntropy=SoftmaxCrossEntropy, LossFunction=LossFunction,
                     decoders=SimpleNamespace(edge=SimpleNamespace(DistMult=DistMult, ComplEx=ComplEx, EdgeDecoder=EdgeDecoder,
                                                                   node_corrupt_forward=node_corrupt_forward, only_pos_forward=only_pos_forward)))
