"""BaseModel: config -> attribute flattening (reference src/base.py:13-41)."""
import inspect
import os


def class_vars(obj):
    return {k: v for k, v in inspect.getmembers(obj)
            if not k.startswith('__') and not callable(v)}


class BaseModel(object):
    def __init__(self, config):
        self.config = config
        self._attrs = class_vars(config)
        for attr in self._attrs:                      # base.py:26-28: strip one leading '_'
            name = attr if not attr.startswith('_') else attr[1:]
            setattr(self, name, getattr(self.config, attr))

    @property
    def checkpoint_dir(self):
        return os.path.join('checkpoints', self.model_dir)

    @property
    def model_dir(self):
        model_dir = self.config.env_name
        for k, v in sorted(self._attrs.items()):
            if not k.startswith('_') and k not in ['display']:
                model_dir += "/%s-%s" % (k, ",".join([str(i) for i in v]) if type(v) == list else v)
        return model_dir + '/'
