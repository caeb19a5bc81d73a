"""NUMA placement helper (aletsch_b200/affinity.py): cpulist parsing and the binding decision, without a GPU."""
import os

from aletsch_b200 import affinity


def test_parse_cpulist():
    assert affinity._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert affinity._parse_cpulist("") == set()


def test_bind_decisions(monkeypatch):
    allowed = os.sched_getaffinity(0)
    calls = []
    monkeypatch.setattr(os, "sched_setaffinity", lambda pid, cpus: calls.append(set(cpus)))
    # no topology: nothing happens
    assert affinity.bind_to_device(0, locality={"bus_id": "x", "numa_node": -1, "cpus": set()})["bound"] is False
    # the device's CPUs lie outside the cpuset
    r = affinity.bind_to_device(0, locality={"bus_id": "x", "numa_node": 1, "cpus": {max(allowed) + 1000}})
    assert r["bound"] is False and "cpuset" in r["why"]
    # every allowed CPU is local
    assert affinity.bind_to_device(0, locality={"bus_id": "x", "numa_node": 0, "cpus": set(allowed)})["bound"] is False
    assert calls == []
    if len(allowed) >= 2:
        half = set(sorted(allowed)[:len(allowed) // 2])
        r = affinity.bind_to_device(0, locality={"bus_id": "x", "numa_node": 0, "cpus": half})
        assert r["bound"] is True and r["cpus"] == len(half) and calls == [half]
    monkeypatch.setenv("AGPU_NO_NUMA_BIND", "1")
    assert affinity.bind_to_device(0, locality={"bus_id": "x", "numa_node": 0, "cpus": {0}})["bound"] is False


def test_device_locality_reads_sysfs(tmp_path, monkeypatch):
    d = tmp_path / "0000:1b:00.0"
    d.mkdir()
    (d / "local_cpulist").write_text("0-15,32-47\n")
    (d / "numa_node").write_text("0\n")
    monkeypatch.setattr(affinity, "_pci_bus_id", lambda device: "00000000:1B:00.0")
    loc = affinity.device_locality(0, sysfs=str(tmp_path))
    assert loc["numa_node"] == 0 and len(loc["cpus"]) == 32 and loc["bus_id"] == "0000:1b:00.0"
    monkeypatch.setattr(affinity, "_pci_bus_id", lambda device: None)
    assert affinity.device_locality(0, sysfs=str(tmp_path)) is None
