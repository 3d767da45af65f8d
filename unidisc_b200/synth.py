"""Synthetic inputs of the shapes SURVEY.md §8(d) fixes for the measured configurations (no datasets in the sandbox)."""
from __future__ import annotations

import torch


def joint_batch(B, txt, img, text_vocab_size, vocab_size, seed):
    """cfg2/cfg3/cfg4: `txt` text tokens followed by `img` image tokens (ids already shifted by text_vocab_size, model.py:200)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.cat([torch.randint(0, text_vocab_size - 1, (B, txt), generator=g),
                     torch.randint(text_vocab_size, vocab_size, (B, img), generator=g)], 1)
    modality = torch.cat([torch.zeros(B, txt, dtype=torch.int64), torch.ones(B, img, dtype=torch.int64)], 1)
    return ids, modality


def packed_batch(B, N, text_vocab_size, vocab_size, seed, min_txt=32, max_txt=512, img_sizes=(256, 1024), tail_pad=True):
    """cfg5 (interleaved / packed, reference dataloader.py:564-678 `PackingCollate` contract): every row is filled with
    documents = (text span of U[min_txt, max_txt] tokens, image of 256 or 1024 tokens) until the next document no longer fits;
    `sample_ids` is the run-length document index, the tail is padding (sample_ids = -1, attention_mask False).
    Returns input_ids, modality, sample_ids, attention_mask and the per-row document lengths."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(B, N, dtype=torch.int64)
    modality = torch.zeros(B, N, dtype=torch.int64)
    sample_ids = torch.full((B, N), -1, dtype=torch.int64)
    doc_lens = []
    for b in range(B):
        pos, doc, lens = 0, 0, []
        while True:
            t = int(torch.randint(min_txt, max_txt + 1, (1,), generator=g))
            im = int(img_sizes[int(torch.randint(0, len(img_sizes), (1,), generator=g))])
            if pos + t + im > N:
                if not tail_pad and N - pos > 4:                 # fill the tail with one last text-only document
                    t = N - pos
                    ids[b, pos:pos + t] = torch.randint(0, text_vocab_size - 1, (t,), generator=g)
                    sample_ids[b, pos:pos + t] = doc
                    lens.append(t)
                break
            ids[b, pos:pos + t] = torch.randint(0, text_vocab_size - 1, (t,), generator=g)
            ids[b, pos + t:pos + t + im] = torch.randint(text_vocab_size, vocab_size, (im,), generator=g)
            modality[b, pos + t:pos + t + im] = 1
            sample_ids[b, pos:pos + t + im] = doc
            lens.append(t + im)
            pos += t + im
            doc += 1
        doc_lens.append(lens)
    attention_mask = sample_ids != -1
    return ids, modality, sample_ids, attention_mask, doc_lens
