"""Hyper-parameter container with the reference's comma-string syntax
(reference util/hparams.py:39-112): ``key=val,key=[a,b],nested=[k=v,k2=v2]``.
Keys and defaults are kept verbatim so the README command lines keep working."""
import copy
import re

_SPLIT_PAIRS = re.compile(r",\s*(?![^\[]*\])")   # commas outside [...]
_SPLIT_KV = re.compile(r"=\s*(?![^\[]*\])")      # '=' outside [...]


class HParams(object):
    def __init__(self, **kwargs):
        self.dict_ = kwargs
        self.__dict__.update(self.dict_)

    def update_config(self, in_string):
        """Update from a comma separated string (hparams.py:45-77): bools by
        ``== 'True'``, nested HParams through ``[k=v,...]``, bracketed lists cast
        to int when the first entry is integral, scalars cast to the default's type."""
        if not in_string:
            return self
        for pair in _SPLIT_PAIRS.split(in_string):
            key, val = _SPLIT_KV.split(pair, maxsplit=1)
            cur = self.dict_[key]
            if isinstance(cur, bool):
                self.dict_[key] = val == "True"
            elif isinstance(cur, HParams):
                cur.update_config(val.strip("[]"))
            else:
                items = val.strip("][").split(",")
                if len(items) > 1:
                    nums = [float(v) for v in items]
                    # the reference tests `val[0].is_integer` without calling it (always truthy),
                    # i.e. every bracketed list becomes a list of ints (hparams.py:69-70)
                    self.dict_[key] = [int(v) for v in nums]
                else:
                    caster = type(cur) if cur is not None else str
                    self.dict_[key] = caster(items[0])
        self.__dict__.update(self.dict_)
        return self

    def save(self, writer):
        for key, val in self.dict_.items():
            if isinstance(val, HParams):
                writer.writerow([key])
                val.save(writer)
            else:
                writer.writerow([key, val])

    def extract(self, keys):
        return {k: self[k] for k in keys if self.exists(k)}

    def keys(self):
        return list(self.dict_)

    def exists(self, key):
        return key in self.dict_

    def __getitem__(self, key):
        return self.dict_[key]

    def __setitem__(self, key, val):
        self.dict_[key] = val
        self.__dict__.update(self.dict_)

    def __deepcopy__(self, memo):
        return HParams(**copy.deepcopy(self.dict_, memo))

    def __repr__(self):
        return "HParams(%s)" % ", ".join(f"{k}={v!r}" for k, v in self.dict_.items())
