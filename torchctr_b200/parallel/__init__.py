from .sharded import ShardedTables, reduce_dense_grads, shard_bases, local_rows
from .model import shard_model
from .peer import IpcTransport, PeerShardedTables, ThreadTransport
from .hybrid import HybridShardedTables

__all__ = ["ShardedTables", "reduce_dense_grads", "shard_bases", "local_rows", "shard_model", "PeerShardedTables", "IpcTransport",
           "ThreadTransport", "HybridShardedTables"]
