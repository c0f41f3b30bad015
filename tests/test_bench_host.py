"""Host-side pieces of bench.py that need no GPU: the NUMA binding of the e2e leg (N > 1) against a made-up sysfs."""
import importlib.util
import io
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


class _Torch:
    class cuda:
        @staticmethod
        def get_device_properties(index):
            return types.SimpleNamespace(pci_domain_id=0, pci_bus_id=0x1B, pci_device_id=0)


def _bench_object(mod):
    o = object.__new__(mod.Bench)
    o.torch, o.local_rank = _Torch, 0
    return o


def test_numa_binding_is_a_quiet_no_op_when_the_topology_is_not_visible(monkeypatch):
    mod = _load_bench()
    before = os.sched_getaffinity(0)
    monkeypatch.delenv("KB200_NUMA", raising=False)
    assert _bench_object(mod).bind_to_gpu_numa_node() is None      # no such PCI function in this container
    assert os.sched_getaffinity(0) == before


def test_numa_binding_restricts_the_rank_to_the_node_of_its_gpu(monkeypatch):
    mod = _load_bench()
    before = os.sched_getaffinity(0)
    if len(before) < 3:
        import pytest
        pytest.skip("needs three CPUs")
    cpus = sorted(before)[:3]
    real_open = open

    def fake_open(path, *a, **k):
        if path == "/sys/bus/pci/devices/0000:1b:00.0/numa_node":
            return io.StringIO("1\n")
        if path == "/sys/devices/system/node/node1/cpulist":
            return io.StringIO(f"{cpus[0]}-{cpus[1]},{cpus[2]},100000\n")   # a CPU outside the allowed set is ignored
        return real_open(path, *a, **k)
    monkeypatch.setattr(mod, "open", fake_open, raising=False)
    monkeypatch.delenv("KB200_NUMA", raising=False)
    try:
        assert _bench_object(mod).bind_to_gpu_numa_node() == 1
        assert os.sched_getaffinity(0) == set(cpus)
        os.sched_setaffinity(0, before)
        monkeypatch.setenv("KB200_NUMA", "0")
        assert _bench_object(mod).bind_to_gpu_numa_node() is None
        assert os.sched_getaffinity(0) == before
    finally:
        os.sched_setaffinity(0, before)
