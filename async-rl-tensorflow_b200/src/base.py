"""What Agent inherits from the reference's model base class (src/base.py:13-41): every data
attribute of the config object becomes an attribute of the model (one leading underscore dropped),
and the checkpoint directory is derived from the public config values.

The directory name keeps the reference's shape -- ``<env_name>/<key>-<value>/.../`` with list values
comma-joined and ``display`` left out -- but the keys are walked in sorted order (the reference
walks a dict whose order depends on the Python version), so that one config always maps to one
directory."""
import os


def config_items(config):
    """name -> value for the data attributes of ``config`` (class or instance): dunder names and
    anything callable are skipped."""
    items = {}
    for name in dir(config):
        if name.startswith('__'):
            continue
        value = getattr(config, name)
        if not callable(value):
            items[name] = value
    return items


def _as_text(value):
    return ",".join(str(x) for x in value) if isinstance(value, list) else str(value)


class BaseModel:
    def __init__(self, config):
        self.config = config
        self._attrs = config_items(config)
        for name, value in self._attrs.items():
            setattr(self, name[1:] if name.startswith('_') else name, value)

    @property
    def model_dir(self):
        parts = [self.config.env_name]
        for key in sorted(self._attrs):
            if key.startswith('_') or key == 'display':
                continue
            parts.append("%s-%s" % (key, _as_text(self._attrs[key])))
        return "/".join(parts) + "/"

    @property
    def checkpoint_dir(self):
        return os.path.join('checkpoints', self.model_dir)
