/*---------------------------------------------------------------------------------------------
 *  NEW FILE for src/vs/workbench/contrib/senweaver/electron-main/apoScoringMainService.ts
 *
 *  Main-process owner of the native addon (napi/apo_napi.c -> libapo_b200.so).  Registered in
 *  src/vs/code/electron-main/app.ts next to the other senweaver channels (app.ts:1258-1274):
 *
 *      services.set(IApoScoringService, new SyncDescriptor(ApoScoringMainService));
 *      mainProcessElectronServer.registerChannel(APO_SCORING_CHANNEL,
 *          ProxyChannel.fromService(accessor.get(IApoScoringService), disposables));
 *
 *  The addon is loaded with a guarded dynamic import, like every optional native module of the
 *  IDE (platform/log/node/spdlogLog.ts:20-35); a failure only disables the fast path.
 *--------------------------------------------------------------------------------------------*/
import { VSBuffer } from '../../../../base/common/buffer.js';
import { ApoScoreBlocks, IApoScoringService } from '../common/apoScoringService.js';

interface ApoAddon {
	create(device: number): unknown;
	rewardBatch(handle: unknown, records: ArrayBuffer): { dims: ArrayBuffer; masks: ArrayBuffer; finals: ArrayBuffer };
	recordsFromJson(utf8: ArrayBuffer): ArrayBuffer;
	score(handle: unknown, dims: ArrayBuffer, C: number, T: number, corpus: ArrayBuffer | null, K: number):
		Promise<{ scores: ArrayBuffer; counts: ArrayBuffer; topk: ArrayBuffer; report: ArrayBuffer }>;
}

function ab(buf: VSBuffer): ArrayBuffer {
	const u8 = buf.buffer;
	return u8.buffer.slice(u8.byteOffset, u8.byteOffset + u8.byteLength) as ArrayBuffer;
}

export class ApoScoringMainService implements IApoScoringService {
	readonly _serviceBrand: undefined;
	private _addon: ApoAddon | undefined;
	private _handle: unknown;
	private readonly _ready: Promise<void>;

	constructor() {
		this._ready = this._load();
	}

	private async _load(): Promise<void> {
		try {
			const mod = await import('apo_b200.node' as string);      // unpacked from the asar by build/gulpfile.vscode.js:315
			this._addon = (mod.default ?? mod) as ApoAddon;
			this._handle = this._addon.create(0);                      // throws when no B200 is visible: no CPU fallback
		} catch (e) {
			console.warn('[APO] native scoring engine unavailable:', e instanceof Error ? e.message : String(e));
			this._addon = undefined;
		}
	}

	async isAvailable(): Promise<boolean> {
		await this._ready;
		return !!this._addon;
	}

	async rewardBatch(records: VSBuffer) {
		await this._ready;
		if (!this._addon) { throw new Error('apo_b200 addon not loaded'); }
		const r = this._addon.rewardBatch(this._handle, ab(records));
		return { dims: VSBuffer.wrap(new Uint8Array(r.dims)), masks: VSBuffer.wrap(new Uint8Array(r.masks)), finals: VSBuffer.wrap(new Uint8Array(r.finals)) };
	}

	async recordsFromJson(persisted: string): Promise<VSBuffer> {
		await this._ready;
		if (!this._addon) { throw new Error('apo_b200 addon not loaded'); }
		return VSBuffer.wrap(new Uint8Array(this._addon.recordsFromJson(ab(VSBuffer.fromString(persisted)))));
	}

	async score(dims: VSBuffer, C: number, T: number, corpus: VSBuffer | undefined, K: number): Promise<ApoScoreBlocks> {
		await this._ready;
		if (!this._addon) { throw new Error('apo_b200 addon not loaded'); }
		// runs on the libuv pool inside the addon (napi_create_async_work): the Electron main loop never blocks
		const r = await this._addon.score(this._handle, ab(dims), C, T, corpus ? ab(corpus) : null, K);
		return {
			scores: VSBuffer.wrap(new Uint8Array(r.scores)), counts: VSBuffer.wrap(new Uint8Array(r.counts)),
			topk: VSBuffer.wrap(new Uint8Array(r.topk)), report: VSBuffer.wrap(new Uint8Array(r.report)),
		};
	}
}
