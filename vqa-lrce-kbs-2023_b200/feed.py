"""Host -> device input feed for the E2E modules. A batch of config-2 clips is 289 MB of fp32 (SURVEY.md §7 hard part 6):
copied on the compute stream it costs ~5 ms in front of a ~20 ms forward. `PrefetchFeed` issues the copy of batch i+1
on a side stream while batch i is being computed, the way a DataLoader with `pin_memory=True` + `non_blocking=True`
is meant to be consumed; the caller's batches must live in pinned host memory for the copy to be asynchronous. The device
side is a ring of two preallocated batch buffers (a yielded batch stays valid until the batch after next is requested)."""
from typing import Iterable, Iterator, Sequence

import torch


class PrefetchFeed:
    """Iterate over host batches (sequences of tensors) and yield them as device tensors, one batch ahead.

        for video_clips, texts, mask, types in PrefetchFeed(loader, device):
            logits = model(video_clips, texts, mask, types)
    """

    def __init__(self, batches: Iterable[Sequence[torch.Tensor]], device: torch.device):
        self.batches, self.device = batches, torch.device(device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._ring = [None, None]  # two device-side batch buffers, reused: no allocator traffic in steady state
        self._n = 0

    def _upload(self, batch):
        slot = self._n & 1
        self._n += 1
        with torch.cuda.stream(self.stream):
            bufs = self._ring[slot]
            if bufs is None or len(bufs) != len(batch) or any(
                    b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, batch)):
                bufs = self._ring[slot] = [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in batch]
            for b, t in zip(bufs, batch):
                b.copy_(t, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return bufs, ev

    def __iter__(self) -> Iterator[Sequence[torch.Tensor]]:
        it = iter(self.batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                # The copy of batch i+1 reuses the buffers of batch i-1: it must not start before everything enqueued on the
                # compute stream so far (the forward of batch i-1 included) has finished reading them.
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
                nxt = self._upload(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            yield cur
