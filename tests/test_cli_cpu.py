"""CPU: flag surface and synthetic dataset of the driver mirror (no compute)."""


def test_options_accept_reference_experiment_flags():
    from dvd_b200.options import options_train
    argv = ('--net scene_flow_motion_field --dataset synthetic_sequence --track_id train --log_time --epoch_batches 2000 '
            '--epoch 20 --lr 1e-6 --html_logger --vali_batches 150 --batch_size 1 --optim adam --vis_batches_vali 4 '
            '--vis_every_vali 1 --vis_every_train 1 --vis_batches_train 5 --vis_at_start --tensorboard --gpu 0 --save_net 1 '
            '--workers 4 --one_way --loss_type l1 --l1_mul 0 --acc_mul 1 --disp_mul 1 --warm_sf 5 --scene_lr_mul 1000 '
            '--repeat 1 --flow_mul 1 --sf_mag_div 100 --time_dependent --gaps 1,2,4,6,8 --midas --use_disp '
            '--logdir ./checkpoints/davis/sequence/ --force_overwrite').split()
    opt, unique = options_train.parse(argv)
    assert opt.midas and opt.use_disp and opt.time_dependent and opt.warm_sf == 5 and opt.scene_lr_mul == 1000
    assert opt.acc_mul == 1 and opt.flow_mul == 1 and opt.disp_mul == 1 and opt.lr == 1e-6
    assert 'gpu' in unique and 'resume' in unique


def test_synthetic_dataset_has_the_reference_batch_keys():
    from dvd_b200.datasets import get_dataset
    from dvd_b200.options import options_train
    opt, _ = options_train.parse(['--net', 'scene_flow_motion_field', '--dataset', 'synthetic_sequence', '--gaps', '1,2,4,6,8',
                                  '--height', '32', '--width', '48'])
    ds = get_dataset('synthetic_sequence')(opt)
    assert len(ds) == 374      # pairs of an 80-frame sequence with gaps 1,2,4,6,8
    b = ds[5]
    for k in ('img_1', 'img_2', 'flow_1_2', 'flow_2_1', 'mask_1', 'mask_2', 'motion_seg_1', 'R_1', 'R_1_T', 'R_2', 'R_2_T',
              't_1', 't_2', 'K', 'K_inv', 'time_stamp_1', 'time_stamp_2', 'frame_id_1', 'frame_id_2', 'time_step'):
        assert k in b, k
    assert b['img_1'].shape == (1, 3, 32, 48) and b['flow_1_2'].shape == (1, 32, 48, 2)


def test_model_rejects_cpu_and_unknown_variants():
    import pytest
    from dvd_b200 import synthetic
    from dvd_b200.models import get_model
    with pytest.raises(NotImplementedError):
        get_model('scene_flow_motion_field')(synthetic.default_opt(use_cnn=True), None)
    m = get_model('scene_flow_motion_field')(synthetic.default_opt(midas=False), None)
    assert [type(n).__name__ for n in m._nets] == ['HourglassModel_Embed', 'SceneFlowFieldNet']
    assert m.num_parameters(return_list=True)[1] == 297987
    batch = synthetic.make_batch([(1, 2)], H=32, W=48)
    with pytest.raises((ValueError, RuntimeError)):
        m._train_on_batch(6, 0, batch)                             # no CPU compute path
