"""tensorflow.python.platform.gfile subset used by the reference's utils/preprocess.py (vocabulary files).
Python 2 read and wrote them as byte strings == str; under Python 3 the same code needs text mode, so the
binary flag is dropped (Latin-1, the encoding of the bundled data files)."""
import os


def Exists(path):
    return os.path.exists(path)


def GFile(path, mode='r'):
    return open(path, mode.replace('b', ''), encoding='latin-1')


Open = GFile
