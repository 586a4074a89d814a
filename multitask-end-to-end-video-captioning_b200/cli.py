"""Entry points behind the reference-named scripts at the repo root (`tf_s2vt.py`, `reinforcement_multisampling_tf_s2vt.py`,
`reinforce_multitask_e2e_attribute_s2vt.py`, `final_beam_search.py`, `e2e_beam_search.py`).

The reference scripts take `--task {train,test,evaluate} --gpu N` and keep every other setting as a module-level
constant edited in place (tf_s2vt.py:275-320).  Here those constants are flags whose defaults are the reference values.
Under `torchrun` (one process per GPU) training is data parallel over NCCL.
"""
import argparse
import json
import os
import random
import time

import numpy as np


def build_parser(description, defaults):
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument('--task', default=defaults.get('task', 'train'), choices=['train', 'test', 'evaluate'], help='tf_s2vt.py:27-37')
    ap.add_argument('--gpu', type=int, default=0, help='device index (tf.device("/gpu:N"), tf_s2vt.py:702); LOCAL_RANK wins under torchrun')
    ap.add_argument('--net', default=None); ap.add_argument('--dataset', default=None)      # parsed and never read by the reference
    ap.add_argument('--tg', default=None); ap.add_argument('--ft', default=None)
    d = dict(video_train_feature_file='data/msvd_train_features.txt', video_test_feature_file='data/msvd_test_features.txt',
             video_train_sent_file='msvd_sents_train_noval_lc_nopunc.txt', video_test_sent_file='msvd_sents_test_lc_nopunc.txt',
             vocabulary_file='msvd_vocabulary1.txt', model_path='models', model_name='s2vt_model', restore=None, out_file='captions.txt',
             dim_image=1536, lstm_dim=1000, word_dim=500, n_video_lstm_step=5, n_caption_lstm_step=35, n_epochs=30, batch_size=64,
             start_learning_rate=1e-3, decay_steps=5000, clip_norm=10.0, dropout_rate=0.9, decay_value=5e-5, seed_num=4,
             n_samples=8, beam_size=3, length_normalization_factor=0.0, alpha=0.5, precision='bf16', max_iters=0,
             reward='cider', beta=10.0, m=0.5, attribute_vocab_file='train_most_freq_vocab_400_truncated.txt')      # beta / m: hinge regulariser of original_attention.py:299-300; reward: cider (cider_evaluation.py) | bleu4 (bleu_evaluation.py) | rouge (rouge_evaluation.py)
    d.update({k: v for k, v in defaults.items() if k != 'task'})
    for k, v in d.items():
        ap.add_argument('--' + k, type=(type(v) if v is not None else str), default=v)
    return ap


def _setup(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', str(args.gpu)))
    torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    rank = dist.get_rank() if world > 1 else 0
    # ONE shuffle stream shared by every rank (the epoch order must be the same permutation everywhere so that the strided batches
    # partition it); rank-dependent randomness lives only in the Philox row offsets of the samplers / dropout (trainer.py).
    random.seed(args.seed_num); np.random.seed(args.seed_num)
    return rank, world


def _vocab_and_model(args, pkg, max_rows_factor=1, beam=1, n_attributes=0, step_name='Variable'):
    vocabulary = pkg.text.read_vocabulary(args.vocabulary_file)
    wordtoix, ixtoword = pkg.text.preProBuildWordVocab(vocabulary, word_count_threshold=0)
    os.makedirs('./new_vocab1_data', exist_ok=True)                                   # tf_s2vt.py:417-419
    np.save('./new_vocab1_data/wordtoix', wordtoix); np.save('./new_vocab1_data/ixtoword', ixtoword)
    model = pkg.Video_Caption_Generator(dim_image=args.dim_image, n_words=len(wordtoix), word_dim=args.word_dim, lstm_dim=args.lstm_dim,
                                        batch_size=args.batch_size, n_lstm_steps=args.n_video_lstm_step + args.n_caption_lstm_step,
                                        n_video_lstm_step=args.n_video_lstm_step, n_caption_lstm_step=args.n_caption_lstm_step,
                                        bias_init_vector=None, decay_value=args.decay_value, dropout_rate=args.dropout_rate,
                                        beam_size=beam, n_attributes=n_attributes, precision=args.precision, max_videos=args.batch_size,
                                        max_rows=args.batch_size * max_rows_factor, seed=args.seed_num)
    model.restored_step = 0
    if args.restore:
        # optimistic_restore (:47-61) runs over tf.global_variables(): the Adam slots and beta powers of the restored variables come
        # along; the step counter only when the checkpoint was written by the same script (`Variable` in tf_s2vt.py, `g_step` in the RL ones)
        restored, model.restored_step = pkg.checkpoint.optimistic_restore(model, args.restore, with_optimizer=True, step_name=step_name)
        print('restored %d variables from %s (global step %d, Adam step %d)' % (len(restored), args.restore, model.restored_step, model.adam_step))
    return wordtoix, ixtoword, model


def _batches(n, bs, rank, world):
    """zip(range(0, n - bs, bs), range(bs, n, bs)) (tf_s2vt.py:482, drops the tail, Q6), strided over ranks.  Every rank gets the
    SAME number of batches (each iteration issues collectives): the up to world-1 surplus batches of an epoch are dropped."""
    starts = list(range(0, n - bs, bs))
    usable = len(starts) // world * world
    return starts[:usable][rank::world]


def run_xe(args):
    """tf_s2vt.py train()/evaluation()/test(): stage-1 cross-entropy S2VT."""
    import s2vt_b200 as pkg
    rank, world = _setup(args)
    wordtoix, ixtoword, model = _vocab_and_model(args, pkg)
    if args.task in ('evaluate', 'test'):
        return _decode_task(args, pkg, model, wordtoix, ixtoword, beam=False)
    train_captions, train_features = pkg.text.get_video_feature_caption_pair(args.video_train_sent_file, args.video_train_feature_file)
    trainer = pkg.trainer.XETrainer(model, args.start_learning_rate, args.decay_steps, args.clip_norm, seed=args.seed_num)
    trainer.global_step = model.restored_step
    it = 0
    for epoch in range(args.n_epochs):
        index = list(range(len(train_captions))); random.shuffle(index)
        for start in _batches(len(index), args.batch_size, rank, world):
            t0 = time.time()
            rows = index[start:start + args.batch_size]
            vids, sents = train_captions[rows, 0], train_captions[rows, 1].tolist()
            feats = np.stack([train_features[v] for v in vids])
            ids, mask = pkg.text.sentence_padding_toix(sents, wordtoix, args.n_caption_lstm_step)
            loss = trainer.step(feats, ids, mask)
            if rank == 0:
                print('idx: ', start, ' Epoch: ', epoch, ' loss: ', float(loss[0].item()), ' Elapsed time: ', str(time.time() - t0))
            it += 1
            if args.max_iters and it >= args.max_iters:
                break
        model.gather_optimizer_state()      # collective: the Adam slots may be sharded over the ranks (trainer.DP_EXCHANGE)
        if rank == 0:
            pkg.checkpoint.save(model, os.path.join(args.model_path, '%s-%d' % (args.model_name, epoch)), trainer.global_step, step_name='Variable')
        if args.max_iters and it >= args.max_iters:
            break


def run_rl(args):
    """reinforcement_multisampling_tf_s2vt.py train(): stage-2 K-sample REINFORCE with CIDEr-D reward and greedy baseline."""
    import s2vt_b200 as pkg
    rank, world = _setup(args)
    wordtoix, ixtoword, model = _vocab_and_model(args, pkg, max_rows_factor=args.n_samples, step_name='g_step')
    if args.task in ('evaluate', 'test'):
        return _decode_task(args, pkg, model, wordtoix, ixtoword, beam=False)
    train_captions, train_features = pkg.text.get_video_feature_caption_pair(args.video_train_sent_file, args.video_train_feature_file)
    by, order = pkg.text.group_by_video(train_captions)
    vindex = {v: i for i, v in enumerate(order)}
    scorer = pkg.rewards.make_scorer(args.reward, [by[v] for v in order], wordtoix)   # CiderD(df=<train corpus>), cider_evaluation.py:12
    trainer = pkg.trainer.ReinforceTrainer(model, scorer, n_samples=args.n_samples, start_learning_rate=args.start_learning_rate,
                                           decay_steps=args.decay_steps, clip_norm=args.clip_norm, seed=args.seed_num)
    trainer.global_step = model.restored_step
    it = 0
    for epoch in range(args.n_epochs):
        index = list(range(len(train_captions))); random.shuffle(index)
        for start in _batches(len(index), args.batch_size, rank, world):
            t0 = time.time()
            vids = train_captions[index[start:start + args.batch_size], 0]             # a batch indexes caption rows (R3)
            feats = np.stack([train_features[v] for v in vids])
            out = trainer.step(feats, np.array([vindex[v] for v in vids], dtype=np.int32))
            if rank == 0:
                print('idx: ', start, ' Epoch: ', epoch, ' loss: ', float(out[1].item()), ' Elapsed time: ', str(time.time() - t0))
            it += 1
            if args.max_iters and it >= args.max_iters:
                break
        model.gather_optimizer_state()      # collective: the Adam slots may be sharded over the ranks (trainer.DP_EXCHANGE)
        if rank == 0:
            pkg.checkpoint.save(model, os.path.join(args.model_path, '%s-%d' % (args.model_name, epoch)), trainer.global_step, step_name='g_step')
        if args.max_iters and it >= args.max_iters:
            break


def run_stage3(args):
    """reinforce_multitask_e2e_attribute_s2vt.py on precomputed features: sum_loss = -(1-lambda) RL + lambda XE (:850),
    lambda = --alpha (0.5), single sample per video, clip 5 (:856)."""
    import torch
    import s2vt_b200 as pkg
    rank, world = _setup(args)
    wordtoix, ixtoword, model = _vocab_and_model(args, pkg, max_rows_factor=1, step_name='g_step')
    if args.task in ('evaluate', 'test'):
        return _decode_task(args, pkg, model, wordtoix, ixtoword, beam=False)
    train_captions, train_features = pkg.text.get_video_feature_caption_pair(args.video_train_sent_file, args.video_train_feature_file)
    by, order = pkg.text.group_by_video(train_captions)
    vindex = {v: i for i, v in enumerate(order)}
    scorer = pkg.cider.CiderD([by[v] for v in order], wordtoix)
    lam, it, step = args.alpha, 0, model.restored_step
    for epoch in range(args.n_epochs):
        index = list(range(len(train_captions))); random.shuffle(index)
        for start in _batches(len(index), args.batch_size, rank, world):
            rows = index[start:start + args.batch_size]
            vids, sents = train_captions[rows, 0], train_captions[rows, 1].tolist()
            feats = torch.from_numpy(np.stack([train_features[v] for v in vids])).to(model.device)
            vi = torch.tensor([vindex[v] for v in vids], dtype=torch.int32, device=model.device)
            gt_ids, gt_mask = pkg.text.sentence_padding_toix(sents, wordtoix, args.n_caption_lstm_step)
            samp, greedy = model.rollout(feats, 1, seed=args.seed_num + step, row_base=rank * len(rows))
            mask, _ = model.caption_masks(samp)
            r = scorer.score_ids(samp, vi).float(); b = scorer.score_ids(greedy, vi).float()
            if world == 1:
                rl = model.rl_backward(feats, samp, mask, r, b, grad_scale=1.0 - lam, drop_seed=step + 1).clone()
                xe = model.xe_backward(feats, gt_ids, gt_mask, grad_scale=lam, accumulate=True, drop_seed=step + 1)
            else:
                # both objectives are normalised by batch-wide mask sums (R1, Q3): exchange those first, then the per-rank shares add up
                import torch.distributed as dist
                gm = torch.from_numpy(np.asarray(gt_mask, dtype=np.float32)).to(model.device)
                stats = torch.cat([mask.sum().view(1), gm.sum(0), torch.tensor([float(len(rows))], device=model.device)])
                dist.all_reduce(stats)
                colsum = stats[1:-1].contiguous()
                rl = model.rl_backward(feats, samp, mask, r, b, norm=float(stats[0].item()), grad_scale=1.0 - lam, drop_seed=step + 1,
                                       row_base=rank * len(rows)).clone()
                xe = model.xe_backward_sharded(feats, gt_ids, gm, colsum, int(round(float(stats[-1].item()))), norm=float(colsum.sum().item()),
                                               decay=(None if rank == 0 else 0.0), grad_scale=lam, accumulate=True, drop_seed=step + 1,
                                               row_base=rank * len(rows))
                pkg.trainer.allreduce_gradients(model, overlap=False)
            lr = pkg.trainer.exponential_decay(args.start_learning_rate, step, args.decay_steps)
            model.optimizer_step(lr, args.clip_norm, wemb_slice_norm=False)
            step += 1; it += 1
            if rank == 0:
                print('idx: ', start, ' Epoch: ', epoch, ' loss: ', float(xe[0].item()))
            if args.max_iters and it >= args.max_iters:
                break
        model.gather_optimizer_state()      # collective: the Adam slots may be sharded over the ranks (trainer.DP_EXCHANGE)
        if rank == 0:
            pkg.checkpoint.save(model, os.path.join(args.model_path, '%s-%d' % (args.model_name, epoch)), step, step_name='g_step')
        if args.max_iters and it >= args.max_iters:
            break


def run_attribute_loss(args):
    """reinforce_multitask_e2e_attribute_loss.py train() on precomputed features: single-sample REINFORCE mixed with the attribute
    head, sum_loss = -(1-alpha) RL + alpha * sigmoidCE(mean_t(X) . attr_W + attr_b, labels) / (400 B) (:375-380, :957); labels from
    get_multilabel over the 400 attribute words (:874-893, :650); clip 10 (:964), Adam 1e-6 halved every 15000 steps (:950-953)."""
    import torch
    import s2vt_b200 as pkg
    rank, world = _setup(args)
    attr_vocab = pkg.text.read_vocabulary(args.attribute_vocab_file)
    wordtoix, ixtoword, model = _vocab_and_model(args, pkg, max_rows_factor=1, n_attributes=len(attr_vocab), step_name='Variable')
    if args.task in ('evaluate', 'test'):
        return _decode_task(args, pkg, model, wordtoix, ixtoword, beam=False)
    train_captions, train_features = pkg.text.get_video_feature_caption_pair(args.video_train_sent_file, args.video_train_feature_file)
    by, order = pkg.text.group_by_video(train_captions)
    vindex = {v: i for i, v in enumerate(order)}
    labels = pkg.text.get_multilabel(by, attr_vocab)
    scorer = pkg.cider.CiderD([by[v] for v in order], wordtoix)
    alpha, it, step = args.alpha, 0, model.restored_step
    for epoch in range(args.n_epochs):
        random.seed(1)                                                               # the reference re-seeds every epoch (:1093)
        index = list(range(len(train_captions))); random.shuffle(index)
        for start in _batches(len(index), args.batch_size, rank, world):
            rows = index[start:start + args.batch_size]
            vids = train_captions[rows, 0]
            feats = torch.from_numpy(np.stack([train_features[v] for v in vids])).to(model.device)
            vi = torch.tensor([vindex[v] for v in vids], dtype=torch.int32, device=model.device)
            y = torch.from_numpy(np.stack([labels[v] for v in vids]).astype(np.float32)).to(model.device)
            samp, greedy = model.rollout(feats, 1, seed=args.seed_num + step, row_base=rank * len(rows))
            mask, _ = model.caption_masks(samp)
            r = scorer.score_ids(samp, vi).float(); b = scorer.score_ids(greedy, vi).float()
            norm = 0.0
            if world > 1:
                import torch.distributed as dist
                tot = mask.sum().view(1).clone(); dist.all_reduce(tot); norm = float(tot.item())
            rl = model.rl_backward(feats, samp, mask, r, b, norm=norm, grad_scale=1.0 - alpha, drop_seed=step + 1, row_base=rank * len(rows)).clone()
            # the head's loss is a mean over the batch rows: a rank's share of the global mean is its local mean / world
            at = model.attribute_backward(feats, y, grad_scale=alpha / world)
            pkg.trainer.allreduce_gradients(model, overlap=False)
            lr = pkg.trainer.exponential_decay(args.start_learning_rate, step, args.decay_steps)
            model.optimizer_step(lr, args.clip_norm, wemb_slice_norm=True)
            step += 1; it += 1
            if rank == 0:
                print('idx: ', start, ' Epoch: ', epoch, ' loss: ', float(rl[0].item()) + alpha * float(at[0].item()))
            if args.max_iters and it >= args.max_iters:
                break
        model.gather_optimizer_state()      # collective: the Adam slots may be sharded over the ranks (trainer.DP_EXCHANGE)
        if rank == 0:
            pkg.checkpoint.save(model, os.path.join(args.model_path, '%s-%d' % (args.model_name, epoch)), step, step_name='Variable')
        if args.max_iters and it >= args.max_iters:
            break


def _decode_task(args, pkg, model, wordtoix, ixtoword, beam):
    """--task test (write captions, tf_s2vt.py:565-609 / final_beam_search.py:504-555) and --task evaluate (greedy or beam
    decode of the test split + CIDEr-D against the test references)."""
    test_captions, test_features = pkg.text.get_video_feature_caption_pair(args.video_test_sent_file, args.video_test_feature_file)
    by, order = pkg.text.group_by_video(test_captions)
    vids = [v for v in test_features]
    t0 = time.time()
    hyps = {}
    for i in range(0, len(vids), args.batch_size):
        chunk = vids[i:i + args.batch_size]
        feats = np.stack([test_features[v] for v in chunk]).astype(np.float32)
        if beam:
            sent, lens, lp, sc = model.beam_search(feats, args.beam_size, args.length_normalization_factor)
            ids = sent.cpu().numpy()
        else:
            ids = model.greedy(feats).cpu().numpy()
        for v, s in zip(chunk, pkg.text.decode_captions(ids, ixtoword)):
            hyps[v] = s
    print('generation time: ', time.time() - t0)
    with open(args.out_file, 'w') as f:
        for v in vids:
            f.write(v + '\t' + hyps[v] + '\n')
    if args.task == 'evaluate':
        keep = [v for v in vids if v in by]
        # evaluate_for_particular_captions / score_all (cider_evaluation.py:14-58): corpus BLEU, ROUGE_L, CIDEr on the GPU scorers
        metrics = pkg.rewards.evaluate_for_particular_captions({v: [hyps[v]] for v in keep}, {v: by[v] for v in keep}, wordtoix)
        score = metrics['CIDEr']
        print(json.dumps(dict(metrics, **{'CIDEr-D': score, 'videos': len(keep)})))
        return score
    return hyps


def run_beam(args):
    """final_beam_search.py / e2e_beam_search.py --task test|evaluate on precomputed features."""
    import s2vt_b200 as pkg
    _setup(args)
    wordtoix, ixtoword, model = _vocab_and_model(args, pkg, beam=args.beam_size)
    return _decode_task(args, pkg, model, wordtoix, ixtoword, beam=True)


def run_attention(args):
    """original_attention.py train() (:390-541) / test() (:544-): temporal-attention decoder -- teacher-forced XE with the hinge
    regulariser on the alphas, Adam 1e-4 halved every 10000 steps, clip 10; per epoch greedy decode of the test split scored with
    evaluate_for_particular_captions; --task test / evaluate decode with the saved alphas available from build_sampler."""
    import s2vt_b200 as pkg
    rank, world = _setup(args)
    train = args.task == 'train'
    if world > 1 and train:
        # the attention handle has no gradient all-reduce (the reference trains it on one GPU, original_attention.py:390-541)
        raise SystemExit('original_attention.py --task train is single-process: launch it without torchrun (WORLD_SIZE=%d)' % world)
    if world > 1 and rank != 0:
        return None                                  # decode tasks: rank 0 alone decodes and writes the output file
    vocabulary = pkg.text.read_vocabulary(args.vocabulary_file)
    wordtoix, ixtoword = pkg.text.preProBuildWordVocab(vocabulary, word_count_threshold=0)
    model = pkg.attention.Video_Caption_Generator(dim_image=args.dim_image, n_words=len(wordtoix), dim_hidden=args.lstm_dim, batch_size=args.batch_size,
                                                  n_video_lstm_steps=args.n_video_lstm_step, n_caption_lstm_steps=args.n_caption_lstm_step,
                                                  drop_out_rate=args.dropout_rate if train else 1.0, bias_init_vector=None, beta=args.beta, m=args.m,
                                                  precision=args.precision, seed=args.seed_num, train=train)
    it0 = 0
    if args.restore:
        restored, it0 = pkg.checkpoint.optimistic_restore(model, args.restore, with_optimizer=True, step_name='Variable')
        print('restored %d variables from %s (global step %d)' % (len(restored), args.restore, it0))

    def decode_test_split():
        test_captions, test_features = pkg.text.get_video_feature_caption_pair(args.video_test_sent_file, args.video_test_feature_file)
        by, _ = pkg.text.group_by_video(test_captions)
        vids, hyps = [v for v in test_features], {}
        for i in range(0, len(vids), args.batch_size):
            chunk = vids[i:i + args.batch_size]
            ids = model.build_generator(np.stack([test_features[v] for v in chunk]).astype(np.float32)).cpu().numpy()
            for v, s_ in zip(chunk, pkg.text.decode_captions(ids, ixtoword)):
                hyps[v] = s_
        keep = [v for v in vids if v in by]
        scores = pkg.rewards.evaluate_for_particular_captions({v: [hyps[v]] for v in keep}, {v: by[v] for v in keep}, wordtoix) if keep else {}
        return vids, hyps, scores

    if not train:
        vids, hyps, scores = decode_test_split()
        with open(args.out_file, 'w') as f:
            for v in vids:
                f.write(v + '\t' + hyps[v] + '\n')
        print(json.dumps(scores))
        return scores if args.task == 'evaluate' else hyps
    train_captions, train_features = pkg.text.get_video_feature_caption_pair(args.video_train_sent_file, args.video_train_feature_file)
    it = 0
    for epoch in range(args.n_epochs):
        index = list(range(len(train_captions))); random.shuffle(index)
        for start in _batches(len(index), args.batch_size, rank, world):
            t0 = time.time()
            rows = index[start:start + args.batch_size]
            vids, sents = train_captions[rows, 0], train_captions[rows, 1].tolist()
            feats = np.stack([train_features[v] for v in vids])
            ids, mask = pkg.text.sentence_padding_toix(sents, wordtoix, args.n_caption_lstm_step)
            loss = model.train_step(feats, ids, mask, global_step=it0 + it, start_learning_rate=args.start_learning_rate, drop_seed=args.seed_num * 7919 + it + 1)
            print('idx: ', start, ' Epoch: ', epoch, ' loss: ', float(loss[0].item()), ' Elapsed time: ', str(time.time() - t0))
            it += 1
            if args.max_iters and it >= args.max_iters:
                break
        _, _, scores = decode_test_split()
        with open(args.out_file, 'a') as f:
            f.write('Epoch %d\n\n' % epoch + ''.join('%s:%s\n' % (k, scores[k]) for k in ('Bleu_1', 'Bleu_2', 'Bleu_3', 'Bleu_4', 'ROUGE_L', 'CIDEr') if k in scores))
        print('CIDEr: ', scores.get('CIDEr'))
        pkg.checkpoint.save(model, os.path.join(args.model_path, 'batch_size%d%s-%d' % (args.batch_size, args.model_name, epoch)), it0 + it, step_name='Variable')
        if args.max_iters and it >= args.max_iters:
            break
