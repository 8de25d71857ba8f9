def is_torch_version(op, version):  # transformer.py:6; only used on the training branch
    return True
