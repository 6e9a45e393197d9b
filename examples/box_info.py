"""What the measurement box gives this process: CPUs (affinity, cgroup quota), memory, NUMA, GPUs / NVLink topology."""
import os, subprocess
print("affinity cpus:", len(os.sched_getaffinity(0)))
for f in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/cpu.stat"):
    try: print(f, "->", open(f).read().strip().replace("\n", " | "))
    except Exception as e: print(f, "->", e)
for cmd in (["nproc"], ["free", "-g"], ["lscpu"], ["nvidia-smi", "topo", "-m"], ["nvidia-smi", "--query-gpu=index,name,memory.total,pcie.link.gen.current,pcie.link.width.current", "--format=csv"]):
    try: print("$", " ".join(cmd)); print(subprocess.run(cmd, capture_output=True, text=True).stdout[:3000])
    except Exception as e: print(e)
