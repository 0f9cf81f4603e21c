"""How many threads the CPU oracle should use on this host (test / CPU-arm infrastructure, like everything under oracle/)."""
import os


def usable_cpus():
    """(logical CPUs this process may run on, physical cores among them, how it was decided). Launchers export
    OMP_NUM_THREADS=1 (torchrun does) — that says nothing about the machine, so it is ignored: the affinity mask, the
    cgroup CPU quota and the sibling lists of /sys decide."""
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cpus = list(range(os.cpu_count() or 1))
    how = f"{len(cpus)} logical CPUs in the affinity mask"
    cores = set()
    for c in cpus:
        try:
            sib = open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list").read().strip()
            pkg = open(f"/sys/devices/system/cpu/cpu{c}/topology/physical_package_id").read().strip()
            cores.add((pkg, sib))
        except OSError:
            cores.add(("?", str(c)))
    n_cores = len(cores)
    quota = None
    try:  # cgroup v2: "max 100000" or "1600000 100000"
        q, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if q != "max":
            quota = max(1, int(q) // int(period))
    except Exception:
        try:  # cgroup v1
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            period = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0:
                quota = max(1, q // period)
        except Exception:
            pass
    if quota is not None and quota < n_cores:
        how += f", {n_cores} physical cores, cgroup quota {quota} CPUs"
        n_cores = quota
    else:
        how += f", {n_cores} physical cores"
    return len(cpus), max(1, n_cores), how


def oracle_threads() -> int:
    """Threads for orc_*_parallel: one per usable physical core. (omp_get_max_threads() answers 128 on a box whose cgroup
    quota is 16 CPUs; 128 threads on 16 CPUs run the OpenMP loops ~25x slower than 16.)"""
    return usable_cpus()[1]
