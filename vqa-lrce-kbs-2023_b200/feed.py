"""Host -> device input feed for the E2E modules. A batch of config-2 clips is 289 MB of fp32 (SURVEY.md §7 hard part 6):
copied on the compute stream it costs ~5 ms in front of a ~20 ms forward. `PrefetchFeed` issues the copy of batch i+1
on a side stream while batch i is being computed, the way a DataLoader with `pin_memory=True` + `non_blocking=True`
is meant to be consumed; the caller's batches must live in pinned host memory for the copy to be asynchronous."""
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

    def _upload(self, batch):
        with torch.cuda.stream(self.stream):
            out = [t.to(self.device, non_blocking=True) for t in batch]
        ev = torch.cuda.Event()
        ev.record(self.stream)
        return out, ev

    def __iter__(self) -> Iterator[Sequence[torch.Tensor]]:
        it = iter(self.batches)
        try:
            nxt = self._upload(next(it))
        except StopIteration:
            return
        while nxt is not None:
            cur, ev = nxt
            try:
                # make sure the copy engine does not overwrite memory the allocator hands back before compute is done
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
                nxt = self._upload(next(it))
            except StopIteration:
                nxt = None
            torch.cuda.current_stream(self.device).wait_event(ev)
            for t in cur:
                t.record_stream(torch.cuda.current_stream(self.device))
            yield cur
