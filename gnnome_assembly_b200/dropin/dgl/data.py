"""dgl.data.DGLDataset as graph_dataset.py:35-68 uses it: the constructor stores the directories and runs
has_cache() / process() / load(); everything else lives in the subclass."""
import os


class DGLDataset:
    def __init__(self, name, url=None, raw_dir=None, save_dir=None, hash_key=(), force_reload=False, verbose=False,
                 transform=None):
        self._name, self._url = name, url
        self._raw_dir = raw_dir if raw_dir is not None else os.path.join(os.path.expanduser("~"), ".dgl")
        self._save_dir = save_dir if save_dir is not None else self._raw_dir
        self._force_reload, self._verbose, self._transform = force_reload, verbose, transform
        self._load()

    # DGLDataset._load: use the cache when there is one, else download + process + save
    def _load(self):
        if not self._force_reload and self.has_cache():
            self.load()
            return
        self.download()
        self.process()
        self.save()

    def download(self):
        pass

    def save(self):
        pass

    def load(self):
        pass

    def process(self):
        raise NotImplementedError

    def has_cache(self):
        return False

    @property
    def name(self):
        return self._name

    @property
    def raw_dir(self):
        return self._raw_dir

    @property
    def raw_path(self):
        return os.path.join(self._raw_dir, self._name)

    @property
    def save_dir(self):
        return self._save_dir

    @property
    def save_path(self):
        return os.path.join(self._save_dir, self._name)

    @property
    def verbose(self):
        return self._verbose
