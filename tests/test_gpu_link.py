"""The drop-in boundary linked against the reference itself: tests/link/link_test.cpp is compiled against the reference's
own headers and linked with oracle/_ref/libmarius_ref.so; a Storage subclass over the C ABI serves the reference's unmodified
GraphModelStorage and DataLoader::loadGPUParameters / updateEmbeddings on cuda:0 in a SynchronousTrainer-shaped loop.  Runs in its own
process (the reference's class names must not meet the product's host adapters in one address space)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "liblink_test.so")


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/liblink_test.so not built (needs /root/reference at build time)")
def test_reference_graph_storage_and_dataloader_over_the_c_abi():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "link", "run_link_test.py")], capture_output=True, text=True, timeout=600)
    assert "LINK_TEST OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
