//! `Transcript`: merlin 3.0.0's `Transcript` with an exportable state.
//!
//! `merlin::Transcript` keeps its STROBE-128 state private, so it cannot cross an FFI boundary.  This type has the same methods and
//! produces the same bytes (checked against merlin's own known-answer test in `tests/test_abi_host.py` through the same C entry
//! points), but its state is the 203-byte wire form of `include/bpp_b200.h` (200 bytes of Keccak state, `pos`, `pos_begin`,
//! `cur_flags`), so a batch of transcripts can be handed to the device, replayed there (`k_replay.cu`) and handed back advanced --
//! exactly what `&mut [Transcript]` means in `RangeProof::verify_batch` (/root/reference/src/range_proof.rs:712-718).
use bpp_b200_sys as sys;

#[derive(Clone)]
pub struct Transcript {
    pub(crate) state: [u8; sys::BPP_TRANSCRIPT_BYTES],
}

impl Transcript {
    /// `merlin::Transcript::new`
    pub fn new(label: &'static [u8]) -> Self {
        let mut state = [0u8; sys::BPP_TRANSCRIPT_BYTES];
        unsafe { sys::bpp_transcript_new(label.as_ptr(), label.len(), state.as_mut_ptr()) };
        Transcript { state }
    }

    /// `merlin::Transcript::append_message`
    pub fn append_message(&mut self, label: &'static [u8], message: &[u8]) {
        unsafe { sys::bpp_transcript_append_message(self.state.as_mut_ptr(), label.as_ptr(), label.len(), message.as_ptr(), message.len()) };
    }

    /// `merlin::Transcript::append_u64`
    pub fn append_u64(&mut self, label: &'static [u8], x: u64) {
        self.append_message(label, &x.to_le_bytes());
    }

    /// `merlin::Transcript::challenge_bytes`
    pub fn challenge_bytes(&mut self, label: &'static [u8], dest: &mut [u8]) {
        unsafe { sys::bpp_transcript_challenge_bytes(self.state.as_mut_ptr(), label.as_ptr(), label.len(), dest.as_mut_ptr(), dest.len()) };
    }

    /// the wire form (what the C ABI exchanges)
    pub fn as_state_bytes(&self) -> &[u8; sys::BPP_TRANSCRIPT_BYTES] {
        &self.state
    }
}
