import torch


def randn_tensor(shape, generator=None, device=None, dtype=None, layout=None):
    """diffusers.utils.torch_utils.randn_tensor: with a CPU generator the draw happens on the CPU
    (in `dtype`) and is then moved to `device` — what tts/utils.py:85 relies on."""
    device = device or torch.device("cpu")
    layout = layout or torch.strided
    rand_device = device
    if generator is not None:
        gen_device_type = generator.device.type if not isinstance(generator, list) else generator[0].device.type
        if gen_device_type != torch.device(device).type and gen_device_type == "cpu":
            rand_device = "cpu"
    return torch.randn(shape, generator=generator, device=rand_device, dtype=dtype, layout=layout).to(device)
