"""``python -m qbot_b200.run FILE`` -- run a qbot program on the CUDA backend; under ``torchrun`` (one process
per GPU, every rank runs the same program) kets of >= QBOT_B200_SHARD_MIN_QUBITS qubits are sharded over the
ranks (qbot_b200/sharded_register.py).  The equivalent of the reference's ``qbot FILE`` (qbot/cli.py:37-56),
which stays the entry point for everything that fits one device."""
import os
import sys


def main(argv=None) -> int:
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m qbot_b200.run FILE", file=sys.stderr)
        return 2
    if int(os.environ.get('WORLD_SIZE', '1')) > 1:
        os.environ.setdefault('QBOT_B200_SHARD', '1')
    import qbot_b200
    from qbot_b200 import sharded_register as sr
    try:
        with open(argv[0]) as f:
            qbot_b200.executeFile(f)
    finally:
        if sr._ctx is not None:
            sr.disable()
            import torch.distributed as dist
            if dist.is_initialized():
                dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
