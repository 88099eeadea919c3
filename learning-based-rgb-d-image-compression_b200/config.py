"""Config object of the reference (utils/IOutils.py:14-22, config/config.py:5-10)."""


class Config(dict):
    """Attribute-style dict, same behaviour as the reference's Config."""

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError as e:
            raise AttributeError(name) from e

    def __setattr__(self, name, value):
        self[name] = value


def model_config():
    return Config({"N": 192, "M": 320, "slice_num": 5, "context_window": 5,
                   "slice_ch": [16, 16, 32, 64, 192], "quant": "ste"})
