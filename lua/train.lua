--[[ train.lua on libdcgansr.so: the host keeps the reference's opt / netG / netD / optimState surface
(/root/reference/train.lua:9-27, 97-152, 275-304); fDx + optim.adam + fGx + optim.adam (:208-283) is one
dcgansr.train_step per iteration.  RGB 32 -> 64 (fineSize 64), MSE family: D regresses 0 on real and the per-sample
pixel MSE / (4 C H W) on fake (:194, :219, :245), G pushes D(fake) to 0 (:264).
Not executed in the build image (no LuaJIT / Torch7 there); see INTEGRATION.md. ]]
require 'torch'
local dsr = require 'dcgansr'
local nn = dsr.nn

opt = {dataset = 'folder', batchSize = 100, loadSize = 96, fineSize = 64, ngf = 12, ndf = 64, nThreads = 4, niter = 1,
       lr = 0.0002, beta1 = 0.5, ntrain = math.huge, name = 'dcgan-sr-test-1', gpu = 1, precision = 'tf32'}
for k, v in pairs(opt) do opt[k] = tonumber(os.getenv(k)) or os.getenv(k) or opt[k] end       -- train.lua:24
print(opt)
opt.manualSeed = torch.random(1, 10000)                                                      -- :30-32
torch.manualSeed(opt.manualSeed)
torch.setdefaulttensortype('torch.FloatTensor')

local DataLoader = paths.dofile('data/data.lua')          -- the reference's own threaded loader, unchanged (:37-38)
local data = DataLoader.new(opt.nThreads, opt.dataset, opt)
print("Dataset: " .. opt.dataset, " Size: ", data:size())

local nc, ndf, ngf = 3, opt.ndf, opt.ngf
local ctx = dsr.Context{gpu = opt.gpu, precision = opt.precision}
local SpatialBatchNormalization, SpatialConvolution, SpatialFullConvolution =
   nn.SpatialBatchNormalization, nn.SpatialConvolution, nn.SpatialFullConvolution

local netG = nn.Sequential()                                                                  -- train.lua:97-113
netG:add(SpatialFullConvolution(nc, ngf * 8, 4, 4, 2, 2, 1, 1))
netG:add(SpatialBatchNormalization(ngf * 8)):add(nn.ReLU(true))
netG:add(SpatialFullConvolution(ngf * 8, ngf * 4, 4, 4, 2, 2, 1, 1))
netG:add(SpatialBatchNormalization(ngf * 4)):add(nn.ReLU(true))
netG:add(SpatialFullConvolution(ngf * 4, ngf * 2, 4, 4, 2, 2, 1, 1))
netG:add(SpatialBatchNormalization(ngf * 2)):add(nn.ReLU(true))
netG:add(SpatialConvolution(ngf * 2, ngf, 4, 4, 2, 2, 1, 1))
netG:add(SpatialBatchNormalization(ngf)):add(nn.LeakyReLU(0.2, true))
netG:add(SpatialConvolution(ngf, nc, 4, 4, 2, 2, 1, 1))
netG:add(nn.Tanh())

local netD = nn.Sequential()                                                                  -- train.lua:119-136
netD:add(SpatialConvolution(nc, ndf, 4, 4, 2, 2, 1, 1))
netD:add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf, ndf * 2, 4, 4, 2, 2, 1, 1))
netD:add(SpatialBatchNormalization(ndf * 2)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 2, ndf * 4, 4, 4, 2, 2, 1, 1))
netD:add(SpatialBatchNormalization(ndf * 4)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 4, ndf * 8, 4, 4, 2, 2, 1, 1))
netD:add(SpatialBatchNormalization(ndf * 8)):add(nn.LeakyReLU(0.2, true))
netD:add(SpatialConvolution(ndf * 8, 1, 4, 4))
netD:add(nn.Sigmoid())
netD:add(nn.View(1):setNumInputDims(3))

netG:cuda(ctx, {nc, opt.fineSize / 2, opt.fineSize / 2}, opt.batchSize)      -- replaces :cuda() / cudnn.convert (:168-180)
netD:cuda(ctx, {nc, opt.fineSize, opt.fineSize}, 2 * opt.batchSize)          -- 2B samples: D(real) and D(fake) run as one grouped pass
dsr.weights_init(netG); dsr.weights_init(netD)                               -- netG:apply(weights_init) (:42-51,114,137)

optimStateG = {learningRate = opt.lr, beta1 = opt.beta1}                     -- :145-152
optimStateD = {learningRate = opt.lr, beta1 = opt.beta1}
-- criterion = nn.MSECriterion(); labels 0 / calMSE / 0  (train.lua:142,219,245,264; calMSE divides by 4*C*H*W, :194)
local step = dsr.StepCfg{criterion = 'MSE', real_label = 0, gen_label = 0, pixel_label = true,
                         pixel_div = 4 * nc * opt.fineSize * opt.fineSize, lr = optimStateD.learningRate, beta1 = optimStateD.beta1}

local epoch_tm, tm, data_tm = torch.Timer(), torch.Timer(), torch.Timer()
for epoch = 1, opt.niter do                                                                   -- train.lua:275-304
   epoch_tm:reset()
   for i = 1, math.min(data:size(), opt.ntrain), opt.batchSize do
      tm:reset()
      data_tm:reset(); data_tm:resume()
      local real_none = data:getBatch()                                                       -- fDx's data:getBatch() (:213)
      data_tm:stop()
      -- (1) Update D network, (2) Update G network with the stale D activations (:280-283)
      local errD_real, errD_fake, errG = dsr.train_step(ctx, netG, netD, step, real_none)
      if ((i - 1) / opt.batchSize) % 1 == 0 then
         print(('Epoch: [%d][%8d / %8d]\t Time: %.3f  DataTime: %.3f    Err_G: %.16f  Err_D: %.4f'):format(
            epoch, ((i - 1) / opt.batchSize), math.floor(math.min(data:size(), opt.ntrain) / opt.batchSize),
            tm:time().real, data_tm:time().real, errG, errD_real + errD_fake))
      end
   end
   print(('End of epoch %d / %d \t Time Taken: %.3f'):format(epoch, opt.niter, epoch_tm:time().real))
end
