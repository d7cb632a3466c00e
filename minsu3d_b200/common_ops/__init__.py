"""Mirror of the reference package layout `minsu3d.common_ops.functions` (boundary #2)."""
