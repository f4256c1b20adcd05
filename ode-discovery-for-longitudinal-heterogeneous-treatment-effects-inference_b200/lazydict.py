"""LazyDict: a dict whose expensive entries are built on first access (shared by dataset.py and counterfactual.py)."""


class LazyDict(dict):
    """The ``.data`` dictionary of a processed dataset.  The three (R, W, k) arrays that nothing on the SINDy / INSITE
    path reads -- one-hot ``current_treatments``, ``prev_treatments``, ``current_covariates``: 3 of the 4 GB a
    10k/1k/1k collection writes -- are built on first access.  Indexing, ``in``, ``get`` see them as ordinary keys;
    anything that enumerates the dictionary (``keys``, ``items``, iteration, ``len``, pickling, ``dict(d)``) builds
    them first, so consumers of the reference's dictionaries cannot tell the difference."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._lazy = {}
        self._on_set = {}
        self.attrs = {}          # side information that travels with copies (e.g. the device-resident compact cohort)
        self._attrs_keys = frozenset()   # assigning one of these keys makes the side information stale: it is dropped

    def set_lazy(self, key, fn):
        dict.pop(self, key, None)
        self._lazy[key] = fn

    def set_lazy_group(self, keys, fn):
        """fn() -> dict with all of `keys`; the first access of any of them runs fn once (one device->host pass) and the
        result is shared by every key and by every copy of this dictionary."""
        cache = {}

        def build(key):
            def one():
                if not cache:
                    cache.update(fn())
                return cache[key]
            return one
        for k in tuple(keys):
            self.set_lazy(k, build(k))

    def pending(self, key):
        """True while `key` has not been built."""
        return key in self._lazy

    def on_set(self, key, fn):
        """fn() is called when `key` is assigned from outside (cached views of it become invalid)."""
        self._on_set[key] = fn

    def materialise(self):
        for k in list(self._lazy):
            self[k]
        return self

    def __missing__(self, key):
        if key not in self._lazy:
            raise KeyError(key)
        v = self._lazy.pop(key)()
        dict.__setitem__(self, key, v)
        return v

    def watch_attrs(self, keys):
        self._attrs_keys = frozenset(keys)

    def __setitem__(self, key, value):
        self._lazy.pop(key, None)
        if key in self._attrs_keys:
            self.attrs.clear()
        if key in self._on_set:
            self._on_set[key]()
        dict.__setitem__(self, key, value)

    def __delitem__(self, key):
        if self._lazy.pop(key, None) is None:
            dict.__delitem__(self, key)

    def __contains__(self, key):
        return dict.__contains__(self, key) or key in self._lazy

    def get(self, key, default=None):
        return self[key] if key in self else default

    def pop(self, key, *default):
        if key in self._lazy:
            self[key]
        return dict.pop(self, key, *default)

    def __iter__(self):
        return dict.__iter__(self.materialise())

    def __len__(self):
        return dict.__len__(self) + len(self._lazy)

    def keys(self):
        return dict.keys(self.materialise())

    def items(self):
        return dict.items(self.materialise())

    def values(self):
        return dict.values(self.materialise())

    def copy(self):
        """Shallow copy that stays lazy (arrays and pending builders are shared)."""
        new = LazyDict()
        dict.update(new, dict.items(self))
        new._lazy = dict(self._lazy)
        new._on_set = dict(self._on_set)   # reassigning a watched key of the copy invalidates the cached views too
        new.attrs = dict(self.attrs)
        new._attrs_keys = self._attrs_keys
        return new

    def __reduce__(self):
        return (dict, (dict(dict.items(self.materialise())),))

    def __deepcopy__(self, memo):
        from copy import deepcopy
        return deepcopy(dict(dict.items(self.materialise())), memo)
