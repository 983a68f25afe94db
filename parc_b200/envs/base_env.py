"""Episode flags shared by the tracker-step operators (reference: envs/base_env.py:12-16)."""
import enum


class DoneFlags(enum.Enum):
    NULL = 0
    FAIL = 1
    SUCC = 2
    TIME = 3
