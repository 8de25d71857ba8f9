import logging as _py_logging


class _Logging:
    ERROR = _py_logging.ERROR
    WARNING = _py_logging.WARNING

    @staticmethod
    def get_logger(name):
        return _py_logging.getLogger(name)


logging = _Logging()


def load_image(x):
    from PIL import Image
    return x if isinstance(x, Image.Image) else Image.open(x)
