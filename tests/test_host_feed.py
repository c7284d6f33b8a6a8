"""ampnet_b200.loader.HostFeed: the double-buffered host -> device feed bench.py's e2e leg uses."""
import pytest
import torch


def test_host_feed_needs_a_cuda_device():
    from ampnet_b200.loader import HostFeed
    with pytest.raises(TypeError):
        HostFeed("cpu")


@pytest.mark.gpu
def test_host_feed_hands_out_batches_in_order_while_the_next_one_travels():
    from ampnet_b200.loader import HostFeed
    dev = torch.device("cuda:0")
    feed = HostFeed(dev)
    n = 1 << 20                                   # every partial sum stays below 2^24: exact in fp32
    hosts = [(torch.full((n,), float(i)).pin_memory(), torch.arange(8, dtype=torch.int64).add_(i).pin_memory())
             for i in range(6)]
    feed.submit(*hosts[0])
    sums = []
    for i in range(6):
        x, idx = feed.get()
        if i + 1 < 6:
            feed.submit(*hosts[i + 1])            # overwrites the buffer set of batch i-1 only after its release()
        y = x
        for _ in range(20):                       # keep the consumer's stream busy while the upload runs
            y = y * 1.0 + 0.0
        sums.append((y.sum(), idx.clone()))
        feed.release()
    torch.cuda.synchronize()
    for i, (s, idx) in enumerate(sums):
        assert float(s.item()) == float(i) * n
        assert idx.cpu().tolist() == list(range(i, i + 8))
    assert feed.bytes_submitted == 6 * (n * 4 + 8 * 8)


@pytest.mark.gpu
def test_host_feed_rejects_pageable_memory_and_overcommit():
    from ampnet_b200.loader import HostFeed
    feed = HostFeed(torch.device("cuda:0"))
    with pytest.raises(TypeError):
        feed.submit(torch.zeros(4))
    a = torch.zeros(4).pin_memory()
    feed.submit(a)
    feed.submit(a)
    with pytest.raises(RuntimeError):
        feed.submit(a)                            # both buffer sets are in flight
    feed.get()
    with pytest.raises(RuntimeError):
        feed.submit(a)                            # one in use, one in flight
    feed.release()
    feed.submit(a)
    with pytest.raises(RuntimeError):
        HostFeed(torch.device("cuda:0")).get()
