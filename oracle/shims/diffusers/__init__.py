from .pipelines import FluxPipeline, DiffusionPipeline  # noqa: F401
