// js/SortWorkerB200.js -- drop-in for src/worker/SortWorker.js of GaussianSplats3D (UNBUILT here: no Node.js in the image).
// Same factory signature and the same message protocol (SortWorker.js:83-113, 202-256); the wasm module and its
// WebAssembly.Memory are replaced by one gs_engine on a B200 through the N-API addon in this directory.
//
// In the reference:   import { createSortWorker } from './worker/SortWorker.js';          (src/Viewer.js:13)
// With this engine:   import { createSortWorker } from 'gsplat-b200/js/SortWorkerB200.js';
import { createRequire } from 'module';
const addon = createRequire(import.meta.url)('./build/Release/gsplat_b200.node');

class B200SortWorker {
    constructor() { this.onmessage = null; this.engine = null; this.uploaded = 0; }

    postMessage(data) {
        if (data.centers) {                                       // SortWorker.js:84-98
            addon.uploadCenters(this.engine, data.centers, this.dynamicMode ? data.sceneIndexes : null, data.range.from, data.range.count);
            this.uploaded = data.range.from + data.range.count;
        } else if (data.sort) {                                   // SortWorker.js:99-113 -> sort() :31-81
            const s = data.sort;
            const renderCount = Math.min(s.splatRenderCount || 0, this.uploaded);
            const sortCount = Math.min(s.splatSortCount || 0, this.uploaded);
            const indexes = this.useSharedMemory ? this.indexesToSort : s.indexesToSort;
            const out = this.useSharedMemory ? this.sortedIndexes : new Uint32Array(renderCount);
            const t0 = performance.now();
            addon.sort(this.engine, new Float32Array(s.modelViewProj), sortCount, renderCount, indexes, out);
            const msg = { sortDone: true, splatSortCount: sortCount, splatRenderCount: renderCount, sortTime: performance.now() - t0 };
            if (!this.useSharedMemory) msg.sortedIndexes = out;
            if (this.onmessage) this.onmessage({ data: msg });
        } else if (data.init) {                                   // SortWorker.js:114-197
            const i = data.init;
            this.useSharedMemory = i.useSharedMemory; this.dynamicMode = i.dynamicMode;
            this.engine = addon.create({ maxSplatCount: i.splatCount, distanceMapRange: i.distanceMapRange,
                                         integerBasedSort: i.integerBasedSort ? 1 : 0, dynamicMode: i.dynamicMode ? 1 : 0 });
            const msg = { sortSetupPhase1Complete: true };
            if (this.useSharedMemory) {
                this.indexesToSort = new Uint32Array(i.splatCount); this.sortedIndexes = new Uint32Array(i.splatCount);
                Object.assign(msg, { indexesToSortBuffer: this.indexesToSort.buffer, indexesToSortOffset: 0,
                                     sortedIndexesBuffer: this.sortedIndexes.buffer, sortedIndexesOffset: 0 });
            }
            queueMicrotask(() => this.onmessage && this.onmessage({ data: msg }));
        }
    }
    terminate() { this.engine = null; }                          // the addon's finalizer calls gs_destroy
}

export function createSortWorker(splatCount, useSharedMemory, enableSIMDInSort, integerBasedSort, dynamicMode,
                                 splatSortDistanceMapPrecision = 16) {
    const worker = new B200SortWorker();
    worker.postMessage({ init: { splatCount, useSharedMemory, integerBasedSort, dynamicMode, distanceMapRange: 1 << splatSortDistanceMapPrecision } });
    return worker;
}
