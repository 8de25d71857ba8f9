from .flux.pipeline_flux import FluxPipeline  # noqa: F401


class DiffusionPipeline:  # only the name is needed by the tts scripts' imports
    @classmethod
    def from_pretrained(cls, *a, **k):
        raise RuntimeError("no checkpoints offline")
