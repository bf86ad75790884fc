"""Write a task's wl_config as a binary blob for non-Python hosts of the C-ABI (examples/c_host).

    python -m wheeledlab_b200.dump_config drift 4096 out.bin [seed]
"""
from __future__ import annotations

import ctypes as C
import sys


def dump(task: str, num_envs: int, path: str, seed: int = 42) -> int:
    from .tasks import make_task
    spec = make_task(task, num_envs=num_envs, seed=seed)
    blob = C.string_at(C.addressof(spec.cfg), C.sizeof(spec.cfg))
    with open(path, "wb") as f:
        f.write(blob)
    return len(blob)


if __name__ == "__main__":
    if len(sys.argv) < 4:
        raise SystemExit(__doc__)
    n = dump(sys.argv[1], int(sys.argv[2]), sys.argv[3], int(sys.argv[4]) if len(sys.argv) > 4 else 42)
    print(f"wrote {n} bytes")
