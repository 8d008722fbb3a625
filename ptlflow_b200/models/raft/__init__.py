from .raft import raft, raft_small  # noqa: F401
