"""Minimal stand-in for the third-party `torchtestcase` package the reference's test-suite is written against
(tests/transforms/transform_test.py:2): tensor-aware assertEqual with an absolute tolerance `eps`."""
import unittest

import torch


class TorchTestCase(unittest.TestCase):
    _eps = None

    @property
    def eps(self):
        return self._eps

    @eps.setter
    def eps(self, value):
        self._eps = value

    def _fail_with_message(self, msg, standard_msg):
        self.fail(self._formatMessage(msg, standard_msg))

    def assertEqual(self, first, second, msg=None):
        if torch.is_tensor(first) and torch.is_tensor(second):
            if first.shape != second.shape:
                self._fail_with_message(msg, "shapes differ: {} vs {}".format(tuple(first.shape), tuple(second.shape)))
            if self._eps is None:
                ok = torch.equal(first, second)
            else:
                ok = first.numel() == 0 or float((first.double() - second.double()).abs().max()) < self._eps
            if not ok:
                self._fail_with_message(msg, "tensors are not equal")
        else:
            super().assertEqual(first, second, msg)

    def assert_tensor_less_equal(self, first, second):
        self.assertTrue(bool((torch.as_tensor(first) <= torch.as_tensor(second)).all()))
