"""Host placement helper of the multi-GPU bench (mocapv2_b200/hostbind.py): sysfs parsing and the no-information fallback."""
import pytest

from mocapv2_b200 import hostbind


def test_cpulist_parsing():
    assert hostbind._cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostbind._cpulist("") == set()
    assert hostbind._cpulist("5") == {5}


def test_unknown_gpu_has_no_node():
    assert hostbind.gpu_numa_node(0xffff, 0xff, 0x1f) == -1


@pytest.mark.gpu
def test_bind_reports_what_it_did():
    import os
    before = os.sched_getaffinity(0)
    rec = hostbind.bind_to_gpu(0)
    try:
        assert set(rec) >= {"numa_node", "cpus_bound", "mempolicy"}
        if rec["numa_node"] < 0:
            assert os.sched_getaffinity(0) == before
        else:
            assert os.sched_getaffinity(0) <= before
    finally:
        os.sched_setaffinity(0, before)
