--[[ train-gray-patch.lua on libdcgansr.so (/root/reference/train-gray-patch.lua:9-25, 54-113, 236-340): one image per
iteration, cut into (fineSize / patchSize)^2 patches of patchSize x patchSize; BCE family with labels 1 / 0 / 1.  The
per-pixel patch loop (:267-273, one CudaTensor write per pixel) becomes dcgansr.stage_patches (one device gather), the
closures + the two optim.adam calls (:236-325, 334-337) one staged dcgansr step.
Not executed in the build image (no LuaJIT / Torch7 there); see INTEGRATION.md. ]]
require 'torch'
require 'image'
local dsr = require 'dcgansr'
local nn = dsr.nn

opt = {batchSize = 64, fineSize = 64, ngf = 16, ndf = 64, niter = 1, lr = 0.0002, beta1 = 0.5, ntrain = 10000, patchSize = 8,
       gpu = 1, precision = 'tf32'}
opt.batchSize = (opt.fineSize / opt.patchSize) * (opt.fineSize / opt.patchSize)              -- :22
for k, v in pairs(opt) do opt[k] = tonumber(os.getenv(k)) or os.getenv(k) or opt[k] end       -- :25
print(opt)
torch.setdefaulttensortype('torch.FloatTensor')
local file_name_route = '/CelebA/Img/img_align_celeba/Img/'
local nc, ndf, ngf = 1, opt.ndf, opt.ngf
local ctx = dsr.Context{gpu = opt.gpu, precision = opt.precision}
local SpatialBatchNormalization, SpatialConvolution, SpatialFullConvolution =
   nn.SpatialBatchNormalization, nn.SpatialConvolution, nn.SpatialFullConvolution

local netG = nn.Sequential()                                                                  -- :54-75
netG:add(nn.SpatialUpSamplingNearest(2))
netG:add(SpatialFullConvolution(nc, ngf * 4, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 4)):add(nn.ReLU(true))
netG:add(SpatialFullConvolution(ngf * 4, ngf * 2, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 2)):add(nn.ReLU(true))
netG:add(SpatialFullConvolution(ngf * 2, ngf, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf, ngf * 2, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 2)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf * 2, ngf * 4, 4, 4, 2, 2, 1, 1)):add(SpatialBatchNormalization(ngf * 4)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf * 4, nc, 4, 4, 2, 2, 1, 1))
netG:add(nn.Sigmoid())

local netD = nn.Sequential()                                                                  -- :94-108 (patch discriminator)
netD:add(SpatialConvolution(nc, ndf, 3, 3)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf, ndf * 2, 3, 3)):add(SpatialBatchNormalization(ndf * 2)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 2, ndf * 4, 3, 3)):add(SpatialBatchNormalization(ndf * 4)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 4, 1, 2, 2))
netD:add(nn.Sigmoid())
netD:add(nn.View(1):setNumInputDims(3))

netG:cuda(ctx, {nc, opt.patchSize / 2, opt.patchSize / 2}, opt.batchSize)
netD:cuda(ctx, {nc, opt.patchSize, opt.patchSize}, 2 * opt.batchSize)
dsr.weights_init(netG); dsr.weights_init(netD)

optimStateG = {learningRate = opt.lr, beta1 = opt.beta1}                                      -- :116-123
optimStateD = {learningRate = opt.lr, beta1 = opt.beta1}
local step = dsr.StepCfg{criterion = 'BCE', real_label = 1, fake_label = 0, gen_label = 1, lr = opt.lr, beta1 = opt.beta1}   -- :113,281,303,320
local line = opt.fineSize / opt.patchSize                     -- patches per image row; the reference's index formula (:270) uses patchSize
                                                              -- here, which is the same number only when fineSize / patchSize == patchSize

local epoch_tm, tm = torch.Timer(), torch.Timer()
for epoch = 1, opt.niter do                                                                   -- :328-340
   epoch_tm:reset()
   local file_num = 1
   for i = 1, opt.ntrain do
      tm:reset()
      local file_name = file_name_route .. ('%06d.jpg'):format(file_num)                      -- :247-259
      local image_input_gray = image.scale(image.load(file_name, 1, 'float'), opt.fineSize, opt.fineSize)
      if image_input_gray:dim() == 3 then image_input_gray = image_input_gray[1] end
      -- real_none[i][a][b] = image[floor((i-1)/line)*patchSize + a][((i-1) % line)*patchSize + b]   (:267-273) on the device
      dsr.stage_patches(ctx, netD, image_input_gray:view(1, opt.fineSize, opt.fineSize):contiguous(), opt.patchSize, line,
                        opt.batchSize, opt.patchSize, 0)
      file_num = file_num + 1
      local errD_real, errD_fake, errG = dsr.train_step_staged(ctx, netG, netD, step, 0, opt.batchSize)   -- :334-337
      print(('errD_real: %.8f  errD_fake: %.8f'):format(errD_real, errD_fake))               -- :309
      print(('Epoch: [%d][%6d / %6d]\t Time: %.3f    Err_G: %.8f  Err_D: %.4f'):format(epoch, i, opt.ntrain, tm:time().real, errG,
                                                                                       errD_real + errD_fake))
   end
   print(('End of epoch %d / %d \t Time Taken: %.3f'):format(epoch, opt.niter, epoch_tm:time().real))
end
