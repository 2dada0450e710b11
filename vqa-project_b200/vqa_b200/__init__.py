"""vqa_b200: host side of the B200-native conditioned-graph VQA hot path.

``layers.py`` / ``sparse_graph_model.py`` next to this package are the drop-in modules; this package holds the
ctypes binding of ``libvqa_sm100.so`` (``_cabi``), tensor-level kernel wrappers (``kernels``), the autograd
operators (``ops``), the data-parallel gradient reducer (``ddp``), the step engine (``engine``), the drivers' criterion and optimiser as
single kernels (``loss``, ``optim``) and the synthetic workload generator.
"""
__all__ = ["_cabi", "kernels", "ops", "ddp", "engine", "loss", "optim", "synthetic"]
