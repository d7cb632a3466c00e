from . import common_ops, hais_ops, pointgroup_ops, softgroup_ops  # noqa: F401
