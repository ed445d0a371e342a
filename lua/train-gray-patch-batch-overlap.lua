--[[ train-gray-patch-batch-overlap.lua on libdcgansr.so (/root/reference/train-gray-patch-batch-overlap.lua:11-26,
76-150, 236-374, 385-694): batchSize images per iteration, each cut into (fineSize / patchSize)^2 non-overlapping training
patches (:262-270); BCE family, labels real_label = 1 / fake_label = 0 / 1.  The training step is that of
train-gray-patch-batch.lua; "overlap" is the test-time path: patches taken every `overlap` pixels (:393-399), generated,
and stitched along minimum-error boundary cuts (:457-694) -- here dcgansr.extract_patches / netG:forward /
dcgansr.stitch_overlap.  (The reference script itself does not run as written: it uses netG / netD / criterion before
defining them, :66-72 vs :76,106,137; the order below is the working one of train-gray-patch-batch.lua.)
Not executed in the build image (no LuaJIT / Torch7 there); see INTEGRATION.md. ]]
require 'torch'
require 'image'
local dsr = require 'dcgansr'
local nn = dsr.nn

opt = {batchSize = 20, fineSize = 64, ngf = 16, ndf = 64, niter = 1, lr = 0.0002, beta1 = 0.5, ntrain = 10000, patchSize = 8,
       overlap = 4, gpu = 1, precision = 'tf32'}
for k, v in pairs(opt) do opt[k] = tonumber(os.getenv(k)) or os.getenv(k) or opt[k] end       -- :25
print(opt)
torch.setdefaulttensortype('torch.FloatTensor')
local patchNumber = (opt.fineSize / opt.patchSize) * (opt.fineSize / opt.patchSize)           -- :28
local file_name_route = '/CelebA/Img/img_align_celeba/Img/'
local nc, ndf, ngf = 1, opt.ndf, opt.ngf
local B = opt.batchSize * patchNumber                                                         -- patches per step (:38-41)
local ctx = dsr.Context{gpu = opt.gpu, precision = opt.precision}
local SpatialBatchNormalization, SpatialConvolution, SpatialFullConvolution =
   nn.SpatialBatchNormalization, nn.SpatialConvolution, nn.SpatialFullConvolution

local netG = nn.Sequential()                                                                  -- :76-102
netG:add(nn.SpatialUpSamplingNearest(2))
netG:add(SpatialFullConvolution(nc, ngf * 4, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 4)):add(nn.ReLU(true))
netG:add(SpatialFullConvolution(ngf * 4, ngf * 2, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 2)):add(nn.ReLU(true))
netG:add(SpatialFullConvolution(ngf * 2, ngf, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf, ngf * 2, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 2)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf * 2, ngf * 4, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 4)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf * 4, nc, 4, 4, 2, 2, 1, 1))
netG:add(nn.Sigmoid())

local netD = nn.Sequential()                                                                  -- :106-122
netD:add(SpatialConvolution(nc, ndf, 3, 3)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf, ndf * 2, 3, 3)):add(SpatialBatchNormalization(ndf * 2)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 2, ndf * 4, 3, 3)):add(SpatialBatchNormalization(ndf * 4)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 4, 1, 2, 2))
netD:add(nn.Sigmoid())
netD:add(nn.View(1):setNumInputDims(3))

netG:cuda(ctx, {nc, opt.patchSize / 2, opt.patchSize / 2}, B)
netD:cuda(ctx, {nc, opt.patchSize, opt.patchSize}, 2 * B)
dsr.weights_init(netG); dsr.weights_init(netD)                                                -- :103,123-135

optimStateG = {learningRate = opt.lr, beta1 = opt.beta1}                                      -- :140-147
optimStateD = {learningRate = opt.lr, beta1 = opt.beta1}
local real_label, fake_label = 1, 0                                                           -- :61-62
local step = dsr.StepCfg{criterion = 'BCE', real_label = real_label, fake_label = fake_label, gen_label = real_label,
                         lr = opt.lr, beta1 = opt.beta1}                                       -- :137,283,305,324
local line = opt.fineSize / opt.patchSize

---- 2. train (:340-374) ----------------------------------------------------------
local images = torch.FloatTensor(opt.batchSize, opt.fineSize, opt.fineSize)
local epoch_tm, tm, data_tm = torch.Timer(), torch.Timer(), torch.Timer()
for epoch = 1, opt.niter do
   epoch_tm:reset()
   local file_set_num = 0
   for i = 1, opt.ntrain, opt.batchSize do
      tm:reset()
      data_tm:reset(); data_tm:resume()
      for k = 1, opt.batchSize do                                                             -- :246-261
         local file_num = file_set_num * opt.batchSize + k
         local img = image.scale(image.load(file_name_route .. ('%06d.jpg'):format(file_num), 1, 'float'), opt.fineSize, opt.fineSize)
         images[k]:copy(img:dim() == 3 and img[1] or img)
      end
      data_tm:stop()
      file_set_num = file_set_num + 1
      dsr.stage_patches(ctx, netD, images, opt.patchSize, line, patchNumber, opt.patchSize, 0)             -- :262-270 on the device
      local errD_real, errD_fake, errG = dsr.train_step_staged(ctx, netG, netD, step, 0, B)                -- :343-346
      print(('Epoch: [%d][%8d / %8d]\t Time: %.3f  DataTime: %.3f    Err_G: %.16f  Err_D: %.4f'):format(
         epoch, ((i - 1) / opt.batchSize) + 1, math.floor(opt.ntrain / opt.batchSize), tm:time().real, data_tm:time().real,
         errG, errD_real + errD_fake))
   end
   print(('End of epoch %d / %d \t Time Taken: %.3f'):format(epoch, opt.niter, epoch_tm:time().real))
end

---- 5. make samples with overlapping patches (:376-694) --------------------------
local real_none_test = image.scale(image.load('/CelebA/Img/img_align_celeba/Img/202001.jpg', 1, 'float'), opt.fineSize, opt.fineSize)
if real_none_test:dim() == 3 then real_none_test = real_none_test[1] end
local overlapPatchLine = (opt.fineSize - opt.overlap) / (opt.patchSize - opt.overlap)         -- :387
local overlapPatchNumber = overlapPatchLine * overlapPatchLine
-- patch i -> rows floor((i-1)/L)*overlap + a, columns ((i-1) % L)*overlap + b   (:393-399)
local real_none_patch_test = dsr.extract_patches(ctx, real_none_test:view(1, opt.fineSize, opt.fineSize):contiguous(), opt.patchSize,
                                                 overlapPatchLine, overlapPatchNumber, opt.overlap)
-- 2 x 2 box down-sample of every patch (:404-410), netG:forward on all of them (:416-418)
local real_reduced = torch.FloatTensor(overlapPatchNumber, 1, opt.patchSize / 2, opt.patchSize / 2)
for i = 1, opt.patchSize / 2 do
   for j = 1, opt.patchSize / 2 do
      real_reduced[{{}, 1, i, j}] = (real_none_patch_test[{{}, 2 * i - 1, 2 * j - 1}] + real_none_patch_test[{{}, 2 * i, 2 * j - 1}] +
                                     real_none_patch_test[{{}, 2 * i - 1, 2 * j}] + real_none_patch_test[{{}, 2 * i, 2 * j}]) / 4
   end
end
local netGt = nn.Sequential()
netGt.layers = netG.layers
netGt:cuda(ctx, {nc, opt.patchSize / 2, opt.patchSize / 2}, overlapPatchNumber)
local p = netG:getParameters()
netGt:setParameters(p)
local fake_patches = netGt:forward(real_reduced)
-- minimum-error boundary cut stitching of the generated patches (:457-694)
local fake_none_test = dsr.stitch_overlap(ctx, fake_patches:view(overlapPatchNumber, opt.patchSize, opt.patchSize), opt.fineSize,
                                          opt.patchSize, opt.overlap, 1)
image.save('fake_none_test.jpg', image.toDisplayTensor(fake_none_test[1]))
print(('PSNR: %.4f  SSIM: %.4f'):format(dsr.calPSNR(ctx, real_none_test:view(1, opt.fineSize, opt.fineSize), fake_none_test)[1],
                                        dsr.calSSIM(ctx, real_none_test:view(1, opt.fineSize, opt.fineSize), fake_none_test)[1]))
