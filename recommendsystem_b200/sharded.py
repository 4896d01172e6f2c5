"""Row-sharded tables + data-parallel dense part across GPUs (placeholder, next commit)."""
