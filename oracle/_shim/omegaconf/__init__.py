"""TEST-ONLY stand-in for the `omegaconf` package.

The reference module gluefactory/models/matchers/lightglue.py imports
`omegaconf` (lightglue.py:9, base_model.py:8-9, utils/experiments.py:14), which
is not installed in this image and cannot be fetched (no network).  This shim
provides just enough surface for the reference file to import and run so that
`oracle/make_golden.py` and `tests/test_oracle_vs_reference.py` can execute the
*unmodified* reference from /root/reference in this container.

It is never imported by product code (glue_factory_colon_b200/ has its own
tiny conf container) and never shipped as a dependency.
"""
from contextlib import contextmanager
from collections.abc import Mapping

__all__ = ["OmegaConf", "DictConfig", "ListConfig", "read_write", "open_dict"]


class DictConfig(dict):
    """dict with attribute access, recursively applied to nested dicts."""

    def __init__(self, content=None):
        super().__init__()
        for k, v in (content or {}).items():
            self[k] = _wrap(v)

    def __getattr__(self, key):
        try:
            return self[key]
        except KeyError as exc:
            raise AttributeError(key) from exc

    def __setattr__(self, key, value):
        self[key] = _wrap(value)

    def get(self, key, default=None):
        return self[key] if key in self else default


class ListConfig(list):
    pass


def _wrap(v):
    if isinstance(v, DictConfig):
        return v
    if isinstance(v, Mapping):
        return DictConfig(v)
    if isinstance(v, (list, tuple)) and not isinstance(v, ListConfig):
        return ListConfig(_wrap(x) for x in v)
    return v


def _unwrap(v):
    if isinstance(v, Mapping):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    return v


def _merge_into(dst, src):
    for k, v in src.items():
        if k in dst and isinstance(dst[k], Mapping) and isinstance(v, Mapping):
            _merge_into(dst[k], v)
        else:
            dst[k] = _wrap(_unwrap(v))


class OmegaConf:
    @staticmethod
    def create(obj=None):
        return DictConfig(_unwrap(obj or {}))

    @staticmethod
    def merge(*confs):
        out = DictConfig()
        for c in confs:
            if c is None:
                continue
            _merge_into(out, c)
        return out

    @staticmethod
    def to_container(conf, resolve=True, **_):
        return _unwrap(conf)

    @staticmethod
    def set_struct(conf, value):
        return None

    @staticmethod
    def set_readonly(conf, value):
        return None

    @staticmethod
    def is_config(obj):
        return isinstance(obj, (DictConfig, ListConfig))

    @staticmethod
    def load(path):
        import yaml

        with open(path) as fh:
            return DictConfig(yaml.safe_load(fh) or {})

    @staticmethod
    def save(conf, path):
        import yaml

        with open(path, "w") as fh:
            yaml.safe_dump(_unwrap(conf), fh)

    @staticmethod
    def from_cli(args=None):
        out = {}
        for a in args or []:
            k, v = a.split("=", 1)
            cur = out
            parts = k.split(".")
            for p in parts[:-1]:
                cur = cur.setdefault(p, {})
            import yaml

            cur[parts[-1]] = yaml.safe_load(v)
        return DictConfig(out)


@contextmanager
def read_write(conf):
    yield conf


@contextmanager
def open_dict(conf):
    yield conf
