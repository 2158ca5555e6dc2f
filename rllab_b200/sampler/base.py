"""Sampler interface (rllab/sampler/base.py:10-38)."""


class Sampler(object):
    def start_worker(self):
        raise NotImplementedError

    def obtain_samples(self, itr):
        raise NotImplementedError

    def process_samples(self, itr, paths):
        raise NotImplementedError

    def shutdown_worker(self):
        raise NotImplementedError
