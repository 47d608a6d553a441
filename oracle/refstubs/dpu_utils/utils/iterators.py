"""Single-threaded equivalents of the dpu_utils iterator helpers (order-preserving)."""
import random


class ThreadedIterator:
    def __init__(self, original_iterator, max_queue_size=2, enabled=True):
        self._it = original_iterator

    def __iter__(self):
        return iter(self._it)


class BufferedIterator(ThreadedIterator):
    pass


def shuffled_iterator(input_iterator, buffer_size=10000, out_slice_sizes=500):
    buf = list(input_iterator)
    random.shuffle(buf)
    yield from buf
