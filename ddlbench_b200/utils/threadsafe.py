"""Condition-variable queue and countdown latch (``pipedream-fork/runtime/threadsafe_queue.py:
9-26``, ``threadsafe_counter.py:9-24``).  The reference hands every received tensor from its
2N helper threads to the main thread through these; this framework's transport pre-posts
receives on the device instead, so they only back host-side producers (the JPEG decode
read-ahead, the launcher's child monitor)."""
from __future__ import annotations

import threading
from collections import deque
from typing import Deque, Generic, Optional, TypeVar

T = TypeVar("T")


class Queue(Generic[T]):
    def __init__(self, maxsize: int = 0):
        self._q: Deque[T] = deque()
        self._cv = threading.Condition()
        self._max = maxsize

    def add(self, item: T) -> None:
        with self._cv:
            while self._max and len(self._q) >= self._max:
                self._cv.wait()
            self._q.append(item)
            self._cv.notify_all()

    def remove(self, timeout: Optional[float] = None) -> T:
        with self._cv:
            if not self._cv.wait_for(lambda: len(self._q) > 0, timeout):
                raise TimeoutError("queue empty")
            item = self._q.popleft()
            self._cv.notify_all()
            return item

    def __len__(self) -> int:
        with self._cv:
            return len(self._q)


class Counter:
    """Starts at ``initial_count``; ``decrement`` by workers, ``wait`` returns at zero."""

    def __init__(self, initial_count: int):
        self._n = initial_count
        self._cv = threading.Condition()

    def decrement(self) -> None:
        with self._cv:
            self._n -= 1
            self._cv.notify_all()

    def wait(self, timeout: Optional[float] = None) -> None:
        with self._cv:
            if not self._cv.wait_for(lambda: self._n <= 0, timeout):
                raise TimeoutError("counter still at %d" % self._n)

    @property
    def count(self) -> int:
        with self._cv:
            return self._n
